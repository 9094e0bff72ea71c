"""Shared helpers for the test-suite (fixture loading, oracle model construction)."""
import json
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MU0 = np.pi * 4.e-7


def load_mesh(name):
    d = np.load(os.path.join(GOLDEN, 'mesh_%s.npz' % name))
    m = dict(r=d['r'], lc=d['lc'], reg=d['reg'], pmap=d['pmap'] if 'pmap' in d else None)
    m['nodesets'] = [d['nodeset%d' % k] for k in range(int(d['n_nodesets']))]
    m['sidesets'] = [d['sideset%d' % k] for k in range(int(d['n_sidesets']))]
    return m


def goldens():
    return json.load(open(os.path.join(GOLDEN, 'goldens.json')))


def split_nodesets(m, jumper_start=0):
    """thincurr_f.F90:172-190: nodesets before `jumper_start` (1-based, negative = from end) are holes."""
    ns = m['nodesets']
    if jumper_start == 0:
        return ns
    js = jumper_start if jumper_start > 0 else len(ns) + 1 + jumper_start
    return ns[:js - 1]


def ref_circle(R, Z, nphi=180):
    """Circular polyline as the reference tests build it (test_ThinCurr.py:330-373)."""
    phi = np.arange(nphi) * (2.0 * np.pi / (nphi - 1))
    return np.stack([R * np.cos(phi), R * np.sin(phi), Z * np.ones(nphi)], 1)


def ref_floop(R, Z, npts=180):
    """circular_flux_loop + save_sensors quantisation ('%.6E', ThinCurr/sensor.py:95-107)."""
    th = np.linspace(0.0, 2.0 * np.pi, npts)
    p = np.stack([R * np.cos(th), R * np.sin(th), Z * np.ones(npts)], 1)
    return np.array([[float('%.6E' % v) for v in row] for row in p])


def dummy_mesh(center, size=1.0, nsplit=0):
    """build_ThinCurr_dummy (ThinCurr/meshing.py:37-85) re-stated for the passive-coil golden."""
    r = np.array([[-size / 2, -size / 2, 0.], [size / 2, -size / 2, 0.], [size / 2, size / 2, 0.],
                  [-size / 2, size / 2, 0.], [0., 0., 0.]]) + np.asarray(center, float)
    lc = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]])
    for _ in range(nsplit):
        lc_new, r_new = [], [x for x in r]
        for j in range(len(lc)):
            ni = [0, 0, 0]
            cand = [(r[lc[j, 0]] + r[lc[j, 1]]) / 2, (r[lc[j, 1]] + r[lc[j, 2]]) / 2, (r[lc[j, 0]] + r[lc[j, 2]]) / 2]
            for k in range(3):
                for k2 in range(len(r), len(r_new)):
                    if np.linalg.norm(r_new[k2] - cand[k]) < 1e-10:
                        ni[k] = k2
                        break
                else:
                    r_new.append(cand[k])
                    ni[k] = len(r_new) - 1
            lc_new += [[lc[j, 0], ni[0], ni[2]], [ni[0], lc[j, 1], ni[1]], [ni[1], lc[j, 2], ni[2]], [ni[0], ni[1], ni[2]]]
        lc, r = np.array(lc_new), np.array(r_new)
    return r, lc


def mutual_abs_sum(O1, O2):
    """A[a][b] = (1/4pi) sum_{c1,c2} sum_comp |E_c1[a]|.|E_c2[b]| T(c1,c2): the magnitude of the terms an
    entry of the mutual matrix is summed from.  |M| << A marks entries dominated by cancellation, where
    the reference's own result moves by ~eps*A with the (atomic, thread-dependent) summation order
    (thin_wall.F90:1092-1121), so parity there is judged against eps*A and not against |M|."""
    import ctypes
    from oracle import tw_oracle as tw
    nc1, nc2 = O1.nc, O2.nc
    P1, P2 = O1.r[O1.lc].reshape(nc1, 9), O2.r[O2.lc].reshape(nc2, 9)
    ii, jj = np.meshgrid(np.arange(nc1), np.arange(nc2), indexing='ij')
    ci, cj = ii.ravel(), jj.ravel()
    Pi, Pj = np.ascontiguousarray(P1[ci]), np.ascontiguousarray(P2[cj])
    Ai, Aj = np.ascontiguousarray(O1.ca[ci]), np.ascontiguousarray(O2.ca[cj])
    n = len(ci)
    To, qo = np.zeros(n), np.zeros(n, np.int32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    tw.lib().tco_pair_T_batch(n, vp(Pi), vp(Ai), vp(Pj), vp(Aj), vp(To), vp(qo))
    T = To.reshape(nc1, nc2)
    E1, E2 = np.zeros((O1.nelems, nc1, 3)), np.zeros((O2.nelems, nc2, 3))
    for c, d in enumerate(O1.cell_basis()):
        for k, v in d.items():
            E1[k, c] = v
    for c, d in enumerate(O2.cell_basis()):
        for k, v in d.items():
            E2[k, c] = v
    A = sum(np.abs(E1[:, :, k]) @ T @ np.abs(E2[:, :, k]).T for k in range(3))
    return A / (4.0 * np.pi)
