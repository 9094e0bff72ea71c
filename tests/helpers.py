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
