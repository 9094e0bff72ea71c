"""Independent pin of the oracle's tw_compute_Bops restatement (thin_wall.F90:1989-2169).  The reference never checks
a B-operator value numerically, so the oracle is checked against physics computed another way:
 * Bel: the defining Biot-Savart surface integral (1/4pi) sum_cells int (e_cell x (r_p - r')) / |r_p - r'|^3 dA'
   evaluated by brute-force subdivision quadrature in numpy (4096 sub-triangles per cell, 3-point rule), for field
   points in the far regime (the reference's quadrature branch) and in the near regime (its finite-difference branch of
   the analytic potential).  Measured agreement: 1e-10...1e-12 beyond ~8 cell sizes (tolerance 1e-8), 5e-9 typically
   and up to 3e-5 at ~2 cell sizes, where the reference's own order selection leaves that much quadrature error
   (tolerance 1e-4: the formula, orientation, scaling and incidence bookkeeping are what this pins);
 * Bdr: the closed form of the midpoint-rule Biot-Savart sum of a regular polygon on its axis (exact to rounding), and
   the smooth-circle formula mu0 I R^2 / (2 (R^2+z^2)^(3/2)) to the polygon's discretisation error;
 * far-field dipole limit of a vertex basis function: B -> (1/4pi) (3 (m.rhat) rhat - m) / r^3 with m = n * (ring area)/3."""
import numpy as np
import pytest
from helpers import MU0, load_mesh, ref_circle
from oracle import tw_oracle as tw


def _subdivide(P, level):
    """Sub-triangles of triangle P[3,3] after `level` uniform 4-way splits: array [4^level, 3, 3]."""
    T = P[None]
    for _ in range(level):
        a, b, c = T[:, 0], T[:, 1], T[:, 2]
        ab, bc, ca = 0.5 * (a + b), 0.5 * (b + c), 0.5 * (c + a)
        T = np.concatenate([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1), np.stack([ca, bc, c], 1), np.stack([ab, bc, ca], 1)])
    return T


def _biot_savart_cell(P, evec, rp, level=6):
    """int_cell (evec x (rp - r')) / |rp - r'|^3 dA' by the 3-point (edge midpoint) rule on 4^level sub-triangles."""
    T = _subdivide(P, level)
    area = 0.5 * np.linalg.norm(np.cross(P[1] - P[0], P[2] - P[0])) / T.shape[0]
    mids = np.concatenate([0.5 * (T[:, 0] + T[:, 1]), 0.5 * (T[:, 1] + T[:, 2]), 0.5 * (T[:, 2] + T[:, 0])])
    d = rp[None] - mids
    r3 = np.linalg.norm(d, axis=1) ** 3
    return np.cross(evec[None], d / r3[:, None]).sum(0) * area / 3.0


@pytest.fixture(scope='module')
def plate():
    m = load_mesh('plate')
    coil = ref_circle(0.35, 0.25, 180)
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], icoils=tw.CoilSets([dict(filaments=[(coil, 1.0, -1.0, -1.0)])]))
    Bel, Bdr = O.compute_Bmat()
    return m, O, Bel, Bdr, coil


def test_bel_against_brute_force_biot_savart(plate):
    m, O, Bel, Bdr, _ = plate
    r, lc = np.asarray(m['r'], float), O.lc
    h = np.sqrt(2.0 * O.ca.mean())
    rng = np.random.default_rng(17)
    dofs = rng.choice(O.np_active, 12, replace=False)
    vert_of = {O.pmap[v] - 1: v for v in range(O.np_) if O.pmap[v] > 0}
    checked_far = checked_near = 0
    for e in dofs:
        ve = vert_of[e]
        cells = [(c, k) for c in range(O.nc) for k in range(3) if lc[c, k] == ve]
        ring = set(lc[[c for c, _ in cells]].ravel())
        dist = np.linalg.norm(r - r[ve], axis=1)
        far = [p for p in np.argsort(dist)[::-1][:3]]                                    # far regime: quadrature branch
        near = [p for p in np.argsort(dist) if p not in ring and dist[p] < 3.0 * h][:3]   # near regime: FD of the potential
        for p in far + near:
            ref = np.zeros(3)
            for c, k in cells:
                ref += _biot_savart_cell(r[lc[c]], O.qbasis[c, k], r[p])
            ref /= 4.0 * np.pi
            got = Bel[:, p, e]
            tol = 1e-8 if dist[p] > 8.0 * h else 1e-4
            assert np.linalg.norm(got - ref) <= tol * np.linalg.norm(ref), (e, p, dist[p] / h, got, ref)
        checked_far += len(far)
        checked_near += len(near)
    assert checked_far >= 30 and checked_near >= 20


def test_bel_far_field_is_a_dipole(plate):
    m, O, Bel, _, _ = plate
    r, lc = np.asarray(m['r'], float), O.lc
    h = np.sqrt(2.0 * O.ca.mean())
    vert_of = {O.pmap[v] - 1: v for v in range(O.np_) if O.pmap[v] > 0}
    worst = 0.0
    for e in (5, 60, 200):
        ve = vert_of[e]
        cells = [c for c in range(O.nc) if ve in lc[c]]
        nrm = np.cross(r[lc[cells[0], 1]] - r[lc[cells[0], 0]], r[lc[cells[0], 2]] - r[lc[cells[0], 0]])
        nrm /= np.linalg.norm(nrm)
        mom = nrm * O.ca[cells].sum() / 3.0       # int psi dA of the hat function, along the oriented normal
        dist = np.linalg.norm(r - r[ve], axis=1)
        for p in np.argsort(dist)[::-1][:5]:
            d = r[p] - r[ve]
            rr = np.linalg.norm(d)
            dip = (3.0 * d * (mom @ d) / rr ** 2 - mom) / rr ** 3 / (4.0 * np.pi)
            err = np.linalg.norm(Bel[:, p, e] - dip) / np.linalg.norm(dip)
            worst = max(worst, err / (h / rr) ** 1)
            assert err < 3.0 * (h / rr), (e, p, err, h / rr)   # next multipole is O(h/r) for the off-centre ring
    assert worst > 0.0


def test_bdr_on_axis_closed_form():
    """Field of the I-coil polyline at a point on its axis: the reference's midpoint Biot-Savart sum
    (thin_wall.F90:2150-2166) has a closed form for a regular polygon."""
    m = load_mesh('plate')
    r = np.asarray(m['r'], float).copy()
    R, Z, n = 0.35, 0.25, 180
    coil = ref_circle(R, Z, n)
    r[0] = [0.0, 0.0, r[0, 2]]   # put one mesh vertex on the coil axis (the operator only reads vertex positions)
    O = tw.OracleModel(r, m['lc'], m['reg'], icoils=tw.CoilSets([dict(filaments=[(coil, 1.0, -1.0, -1.0)])]))
    _, Bdr = O.compute_Bmat()
    z = r[0, 2] - Z
    nseg = n - 1                                     # ref_circle closes the loop: n points, n-1 segments
    L, Rm = 2.0 * R * np.sin(np.pi / nseg), R * np.cos(np.pi / nseg)
    bz = MU0 / (4.0 * np.pi) * nseg * L * Rm / (Rm ** 2 + z ** 2) ** 1.5
    got = Bdr[:, 0, 0]
    assert abs(got[2] - bz) <= 1e-12 * abs(bz) and np.abs(got[:2]).max() <= 1e-12 * abs(bz)
    smooth = MU0 * R ** 2 / (2.0 * (R ** 2 + z ** 2) ** 1.5)
    assert abs(got[2] / smooth - 1.0) < 2.0 * (np.pi / nseg) ** 2
