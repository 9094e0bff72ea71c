"""CPU checks of the oracle's restatements of the SURVEY 8f rows (tw_compute_Lmatblock / -LmatHole / -Bops_block,
thin_wall_hodlr.F90:136-404,580-691; tw_compute_Lmat_MF, thin_wall.F90:1190-1414) against the parts of the oracle that the
reference's goldens pin (tests/test_oracle_golden.py): the block builders must reproduce the entries of the dense
operators up to the quadrature error of the swapped near-field role, the B blocks exactly (same formula), and the
matrix-free apply the mutual matrix times a vector up to its coarser 3-level quadrature.  No reference test reads these
routines' outputs entry-wise (they are exercised through test_ThinCurr.py's HODLR cases :1229-1245 only), so beyond this
they are 'parity unpinned', like L itself."""
import numpy as np
from helpers import load_mesh, split_nodesets
from oracle import tw_oracle as tw


def model(name, jumper_start=0):
    m = load_mesh(name)
    ns = split_nodesets(m, jumper_start)
    cl = m['sidesets'][0] if m['sidesets'] else ()
    return tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=ns, closures=cl)


def test_lmat_block_matches_dense_entries():
    O = model('plate')
    L = O.compute_Lmat()
    act = np.nonzero(O.pmap > 0)[0].astype(np.int32)
    rng = np.random.default_rng(5)
    rows, cols = rng.permutation(act)[:60], rng.permutation(act)[:90]
    B = O.lmat_block(rows, cols)
    ref = L[np.ix_(O.pmap[rows] - 1, O.pmap[cols] - 1)]
    # same far field; near pairs differ by which triangle is integrated analytically (row block's here, smaller DOF's there)
    assert np.abs(B - ref).max() < 1e-5 * np.abs(L).max()
    # one-row strips (the ACA access pattern, thin_wall_hodlr.F90:1260-1283) are rows of the block
    S = O.lmat_block(rows[:1], cols)
    assert np.abs(S[0] - B[0]).max() < 1e-13 * np.abs(B).max()


def test_lmat_block_role_is_the_row_block():
    """Swapping the blocks swaps the analytic side: B(rows,cols) != B(cols,rows)^T in the near field, in the last digits
    of the quadrature error only; identical where no pair is near (disjoint far blocks)."""
    O = model('plate')
    x = O.r[:, 0]
    act = np.nonzero(O.pmap > 0)[0]
    left, right = act[x[act] < np.quantile(x[act], 0.2)].astype(np.int32), act[x[act] > np.quantile(x[act], 0.8)].astype(np.int32)
    A, B = O.lmat_block(left, right), O.lmat_block(right, left)
    assert np.abs(A - B.T).max() < 1e-10 * np.abs(A).max()  # summation order of the swapped far sums


def test_lmat_hole_matches_dense_columns():
    O = model('torus')
    assert O.nholes >= 1
    L = O.compute_Lmat()
    H = O.lmat_hole()
    assert H.shape == (O.nholes, O.nelems)
    ref = L[O.np_active:O.np_active + O.nholes, :]
    assert np.abs(H - ref).max() < 1e-5 * np.abs(L).max()


def test_bops_block_matches_dense_operator():
    O = model('cyl', jumper_start=2)
    Bel, _ = O.compute_Bmat()
    act = np.nonzero(O.pmap > 0)[0].astype(np.int32)
    rng = np.random.default_rng(7)
    rows = rng.permutation(act)[:40]
    cols = rng.permutation(O.np_)[:70].astype(np.int32)
    for d in range(3):
        B = O.bops_block(rows, cols, d)
        ref = Bel[d][np.ix_(cols, O.pmap[rows] - 1)].T
        assert np.abs(B - ref).max() < 1e-12 * np.abs(Bel).max()


def test_cross_eval_matches_mutual_matrix():
    O1, O2 = model('plate'), model('cyl', jumper_start=2)
    O2.r[:, 2] += 0.0  # (same frame: the plate sits inside the cylinder's bore)
    M = O1.cross_coupling(O2)  # [nelems1][nelems2]
    rng = np.random.default_rng(11)
    a = rng.standard_normal((3, O1.nelems))
    counts = np.zeros(3, np.int64)
    b = O1.cross_eval(O2, a, counts)
    assert counts.sum() == O1.nc * O2.nc
    ref = a @ M
    assert np.abs(b - ref).max() < 1e-4 * np.abs(ref).max()
    # linear in the right-hand sides
    b2 = O1.cross_eval(O2, a[:1] * 2.0 - a[1:2])
    assert np.abs(b2[0] - (2.0 * b[0] - b[1])).max() < 1e-12 * np.abs(b).max()


def test_cross_eval_self_has_all_classes():
    O = model('plate')
    rng = np.random.default_rng(13)
    a = rng.standard_normal((1, O.nelems))
    counts = np.zeros(3, np.int64)
    b = O.cross_eval(O, a, counts)
    assert (counts > 0).all()
    L = O.compute_Lmat()
    ref = a @ L
    assert np.abs(b - ref).max() < 5e-3 * np.abs(ref).max()
