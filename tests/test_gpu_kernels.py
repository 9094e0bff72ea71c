"""Kernel-level GPU parity: the device pair integral T(i,j), the order selection and the analytic
potential, evaluated by the same device functions the operator kernels use, against the oracle's
tco_pair_T / tco_phipot (thin_wall.F90:1044-1083, :1934-1985) on seeded random pairs of the
reference's own test meshes.  The probes live in the TEST build of the library
(libthincurr_b200_test.so, -DTW_TEST_HOOKS: same sources, same device functions), not in the product library.  Tolerances: order selection bit-exact (integer); T within 2e-14
relative (far field differs from the CPU only by summation order / FMA; the near field is evaluated
with the reference's operation order and IEEE sqrt/div, differing by the libm log/atan2 only)."""
import ctypes
import numpy as np
import pytest
from helpers import load_mesh, split_nodesets
from oracle import tw_oracle as tw

pytestmark = pytest.mark.gpu


def _pairs(name, js, n, seed):
    m = load_mesh(name)
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=split_nodesets(m, js),
                       closures=m['sidesets'][0] if m['sidesets'] else ())
    rng = np.random.default_rng(seed)
    ci, cj = rng.integers(0, O.nc, n), rng.integers(0, O.nc, n)
    # force a good share of self / adjacent pairs (order 18, shared vertices)
    cj[: n // 20] = ci[: n // 20]
    P = O.r[O.lc].reshape(O.nc, 9)
    return (np.ascontiguousarray(P[ci]), np.ascontiguousarray(O.ca[ci]), np.ascontiguousarray(P[cj]), np.ascontiguousarray(O.ca[cj]))


def _oracle_T(Pi, Ai, Pj, Aj):
    n = len(Ai)
    T, q = np.zeros(n), np.zeros(n, np.int32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    tw.lib().tco_pair_T_batch(n, vp(Pi), vp(Ai), vp(Pj), vp(Aj), vp(T), vp(q))
    return T, q


@pytest.mark.parametrize('name,js', [('plate', 0), ('cyl', 2), ('torus', 0), ('ex_torus', 0)])
@pytest.mark.parametrize('mode', [0, 1])
def test_pair_integrals(name, js, mode):
    from openfusiontoolkit_b200 import _testlib as I
    Pi, Ai, Pj, Aj = _pairs(name, js, 200000, 7)
    To, qo = _oracle_T(Pi, Ai, Pj, Aj)
    Tg, qg = np.zeros_like(To), np.zeros_like(qo)
    assert I.b200_probe_pairs(len(Ai), mode, Pi, Ai, Pj, Aj, Tg, qg) == 0
    assert np.array_equal(qg & 31, qo), 'quadrature order must match the reference expression bit for bit'
    rel = np.abs(Tg - To) / np.abs(To)
    assert rel.max() < 2e-14, 'max rel err of T %.3e (order %d)' % (rel.max(), qo[np.argmax(rel)])
    if mode == 1:
        assert ((qg & 64) != 0).mean() < 0.01, 'FP32 screen should settle almost every pair'


def test_pair_integrals_scaled_geometry():
    """Order selection / evaluation must not depend on the length scale or the position in space."""
    from openfusiontoolkit_b200 import _testlib as I
    Pi, Ai, Pj, Aj = _pairs('torus', 0, 50000, 11)
    for scale, shift in ((1e-3, 0.0), (37.0, 0.0), (1.0, 1000.0)):
        Pi2, Pj2 = Pi * scale + shift, Pj * scale + shift
        Ai2, Aj2 = Ai * scale ** 2, Aj * scale ** 2
        To, qo = _oracle_T(Pi2, Ai2, Pj2, Aj2)
        Tg, qg = np.zeros_like(To), np.zeros_like(qo)
        assert I.b200_probe_pairs(len(Ai), 1, Pi2, Ai2, Pj2, Aj2, Tg, qg) == 0
        assert np.array_equal(qg & 31, qo)
        # a far-away origin costs digits in BOTH implementations (cancellation in the vertex differences)
        rel = np.abs(Tg - To) / np.abs(To)
        if shift == 0.0:
            # far field: a few ulp; analytic near field: libm log/atan2 differences amplified by the
            # conditioning of the potential formula (thin_wall.F90:1966-1983)
            far = qo <= 10
            assert rel[far].max() < 2e-14 and rel[~far].max() < 2e-13, (rel[far].max(), rel[~far].max())
        else:
            assert rel.max() < 1e-9


def test_phipot():
    from openfusiontoolkit_b200 import _testlib as I
    rng = np.random.default_rng(3)
    n = 100000
    tri = rng.normal(size=(n, 9))
    pt = rng.normal(size=(n, 3)) * 2.0
    # points in the plane of the triangle, on edges' extensions and at vertices (guards of :1966-1970)
    t3 = tri.reshape(n, 3, 3)
    pt[:1000] = t3[:1000, 0] + 1.7 * (t3[:1000, 1] - t3[:1000, 0])
    pt[1000:2000] = t3[1000:2000, 2]
    pt[2000:3000] = (t3[2000:3000, 0] + t3[2000:3000, 1] + t3[2000:3000, 2]) / 3.0
    out_g, out_o = np.zeros(n), np.zeros(n)
    assert I.b200_probe_phipot(n, np.ascontiguousarray(tri), np.ascontiguousarray(pt), out_g) == 0
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    tw.lib().tco_phipot_batch(n, vp(tri), vp(pt), vp(out_o))
    ok = np.isfinite(out_o)
    assert np.array_equal(np.isfinite(out_g), ok)
    err = np.abs(out_g[ok] - out_o[ok]) / np.maximum(np.abs(out_o[ok]), 1e-3)
    assert err.max() < 1e-12


def test_rsqrt():
    from openfusiontoolkit_b200 import _testlib as I
    x = np.exp(np.random.default_rng(0).uniform(-60, 60, 300000))
    y = np.zeros_like(x)
    assert I.b200_probe_rsqrt(len(x), x, y) == 0
    assert np.abs(y * np.sqrt(x) - 1.0).max() < 4e-16
