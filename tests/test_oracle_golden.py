"""Pin the CPU oracle (oracle/) against every golden the reference holds for this path.

Goldens quoted from /root/reference/src/tests/physics/test_ThinCurr.py (eigenvalues :984,:1029,
:1076,:1164, tol 1e-5 at :436; frequency response :1004,:1050,:1099,:1189, tol 1e-4 at :464) and
the quadrature KAT src/tests/grid/quad_2d.tests (tol 1e-12, test_quad.py:49-60).
"""
import numpy as np
import pytest
from helpers import (MU0, dummy_mesh, goldens, load_mesh, split_nodesets, ref_circle, ref_floop)
from oracle import tw_oracle as tw

G = goldens()


def _model(name, g, vcoils=None, icoils=None):
    if name == 'passive':
        r, lc = dummy_mesh([0.0, 0.0, 10.0], size=0.25, nsplit=1)  # run_eig, test_ThinCurr.py:186-187
        m = dict(r=r, lc=lc, reg=None, nodesets=[], sidesets=[])
    else:
        m = load_mesh(name)
    eta = g.get('eta', 10.0) * MU0
    vc = tw.CoilSets([dict(filaments=[(ref_circle(R, Z), 1.0, 1.e-2, 1.256637E-5)]) for (R, Z) in (vcoils or [])])
    ic = tw.CoilSets([dict(filaments=[(ref_circle(R, Z), 1.0, -1.0, -1.0) for (R, Z) in icoils])] if icoils else [])
    return tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=split_nodesets(m, g.get('jumper_start', 0)),
                          closures=m['sidesets'][0] if m['sidesets'] else (), eta=[eta], vcoils=vc, icoils=ic)


@pytest.mark.parametrize('name', ['plate', 'cyl', 'torus', 'passive'])
def test_eigenvalue_goldens(name):
    g = G['eig_' + name]
    M = _model(name, g, vcoils=g.get('vcoils'))
    M.compute_Mcoil()
    M.compute_Lmat()
    M.compute_Rmat()
    e = M.get_eigs(4)
    assert np.abs(e / np.array(g['vals']) - 1.0).max() < g['tol']


@pytest.mark.parametrize('name', ['plate', 'cyl', 'torus', 'passive'])
def test_frequency_response_goldens(name):
    """Pins Mcoil + Msensor + L + R jointly: (i w L + R) x = -i w M I (thin_wall_solvers.F90:279-309)."""
    g = G['fr_' + name]
    M = _model(name, g, vcoils=g.get('vcoils'), icoils=g['icoils'])
    Mc = M.compute_Mcoil()
    Ms, Msc = M.compute_Msensor([(ref_floop(R, Z), 1.0) for (R, Z) in g['floops']])
    L = M.compute_Lmat()
    R = M.compute_Rmat().toarray()
    dc = 1.0 / MU0
    om = 2.0 * np.pi * g['freq']
    b = 1j * (-om * Mc[0] * dc)
    x = np.linalg.solve(1j * om * L + R, b)
    sig = np.stack([x.real, x.imag]) @ Ms
    sig[0] += dc * Msc[0]
    assert np.abs(sig[0] / np.array(g['real']) - 1.0).max() < g['tol']
    assert np.abs(sig[1] / np.array(g['imag']) - 1.0).max() < g['tol']


def test_quadrature_kat():
    """Monomial integrals over the unit square split in two triangles (test_quad.F90:55-68)."""
    import ctypes
    import os
    kat = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'quad_2d_kat.npy'))
    lib = tw.lib()
    tris = [np.array([[0, 0], [1, 0], [1, 1.]]), np.array([[0, 0], [1, 1], [0, 1.]])]
    worst = 0.0
    for e1, e2, c1, c2, y0 in kat:
        order = int(max(e1, e2, 1))
        if order > 18:
            continue
        n = lib.tco_quad_np(order)
        pts = np.zeros((n, 3))
        wts = np.zeros(n)
        lib.tco_quad_get(order, pts.ctypes.data_as(ctypes.c_void_p), wts.ctypes.data_as(ctypes.c_void_p))
        assert abs(wts.sum() - 1.0) < 1e-14 and np.abs(pts.sum(1) - 1.0).max() < 1e-15
        y = 0.0
        for T in tris:
            x = pts @ T
            y += 0.5 * (wts * (c1 * x[:, 0] ** e1 + c2 * x[:, 1] ** e2)).sum()
        worst = max(worst, abs((y - y0) / y0))
    assert worst < 1e-12


def test_role_asymmetry_and_far_symmetry():
    """SURVEY hard part 1: near T(i,j) != T(j,i); far pairs symmetric to rounding."""
    import ctypes
    M = _model('plate', G['eig_plate'])
    lib = tw.lib()
    P = M.r[M.lc]
    A3 = ctypes.c_double * 9

    def T(i, j):
        iq = ctypes.c_int()
        v = lib.tco_pair_T(A3(*P[i].ravel()), ctypes.c_double(M.ca[i]), A3(*P[j].ravel()), ctypes.c_double(M.ca[j]), ctypes.byref(iq))
        return v, iq.value
    near = far = 0.0
    for i in range(0, 60):
        for j in range(i + 1, M.nc, 7):
            a, q = T(i, j)
            b, _ = T(j, i)
            d = abs(a - b) / abs(a)
            if q > 10:
                near = max(near, d)
            else:
                far = max(far, d)
    assert far < 1e-12 and 1e-9 < near < 1e-3


def test_simple_hash_known_answers():
    """Jenkins one-at-a-time (oft_local_c.c:86-98) known answers for ASCII keys."""
    import ctypes
    lib = tw.lib()
    for key, want in ((b'a', 0xca2e9442), (b'The quick brown fox jumps over the lazy dog', 0x519e91f5)):
        buf = ctypes.create_string_buffer(key, len(key))
        got = lib.tco_simple_hash(ctypes.addressof(buf), len(key)) & 0xffffffff
        assert got == want
