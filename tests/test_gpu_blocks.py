"""GPU parity of the SURVEY 8f rows built on the cell-list sweeps (csrc/tw_blocks.cu) against the oracle:
tw_compute_Lmatblock / tw_compute_LmatHole / tw_compute_Bops_block (thin_wall_hodlr.F90:136-404,580-691),
tw_compute_Lmat_MF behind ThinCurr.cross_eval (thin_wall.F90:1190-1414) and the projections of tw_reduce_model
(thin_wall_solvers.F90:1180-1359).  Tolerances: entries 1e-10 relative to themselves where they are not small against
the largest entry of the block (|x| > 1e-6 max), 1e-10 of the largest entry elsewhere (sums with cancellation)."""
import os
import numpy as np
import pytest
from helpers import MU0, load_mesh, mutual_abs_sum, split_nodesets, ref_circle, ref_floop
from oracle import tw_oracle as tw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    from openfusiontoolkit_b200 import OFT_env
    return OFT_env(nthreads=-1)


def make(env, name, jumper_start=0, vcoils=None, icoils=None):
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh(name)
    ns = split_nodesets(m, jumper_start)
    cl = m['sidesets'][0] if m['sidesets'] else None
    vc = [[dict(pts=ref_circle(R, Z), scale=1.0, radius=1.e-2, res_per_len=1.256637E-5)] for (R, Z) in (vcoils or [])]
    ic = [[dict(pts=ref_circle(R, Z), scale=1.0) for (R, Z) in icoils]] if icoils else []
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=ns, closures=cl if cl is not None else (), eta=[10.0 * MU0],
                       vcoils=tw.CoilSets([dict(filaments=[(f['pts'], 1.0, 1.e-2, 1.256637E-5) for f in s]) for s in vc]),
                       icoils=tw.CoilSets([dict(filaments=[(f['pts'], 1.0, -1.0, -1.0) for f in s]) for s in ic]))
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=ns if len(ns) else None, closures=cl)
    if vc:
        T.set_coils('vcoil', vc)
    if ic:
        T.set_coils('icoil', ic)
    T.set_eta_values(eta_surf=np.array([10.0 * MU0]))
    return O, T


def close(A, B, tol=1e-10, small=1e-6):
    A, B = np.asarray(A), np.asarray(B)
    assert A.shape == B.shape
    scale = np.abs(B).max()
    big = np.abs(B) > small * scale
    e1 = (np.abs(A - B)[big] / np.abs(B)[big]).max() if big.any() else 0.0
    e2 = (np.abs(A - B)[~big]).max() / scale if (~big).any() else 0.0
    assert e1 < tol and e2 < tol, (e1, e2)
    return max(e1, e2)


@pytest.mark.parametrize('name,js', [('plate', 0), ('cyl', 2), ('torus', 0)])
def test_lmatblock_vs_oracle(env, name, js):
    O, T = make(env, name, js)
    rng = np.random.default_rng(3)
    pts = rng.permutation(O.np_).astype(np.int32)
    rows, cols = np.sort(pts[:150]), pts[100:420]  # overlapping blocks: shared cells, coincident pairs (i == j)
    Bo = O.lmat_block(rows, cols)
    Bg = T.compute_Lmatblock(rows, cols)
    close(Bg, Bo)
    # one-row strips against a column block (ACA+ access pattern, thin_wall_hodlr.F90:1224-1283), host and device output
    import torch
    for k in (0, 17, 149):
        So = O.lmat_block(rows[k:k + 1], cols)
        Sg = T.compute_Lmatblock(rows[k:k + 1], cols)
        close(Sg, So)
        d = torch.zeros((1, len(cols) + 3), dtype=torch.float64, device='cuda')
        T.compute_Lmatblock(rows[k:k + 1], cols, out=d)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy()[:, :len(cols)], Sg)
    # deterministic: same call, same bits
    assert np.array_equal(T.compute_Lmatblock(rows, cols), Bg)


def test_lmatblock_two_models(env):
    """row and column blocks from different models (tw_compute_Lmatblock(row_obj, col_obj, ...))"""
    O1, T1 = make(env, 'plate')
    O2, T2 = make(env, 'cyl', 2)
    rows = np.nonzero(O1.pmap > 0)[0][::3].astype(np.int32)
    cols = np.nonzero(O2.pmap > 0)[0][::2].astype(np.int32)
    Bg, Bo = T1.compute_Lmatblock(rows, cols, col_model=T2), O1.lmat_block(rows, cols, other=O2)
    # entries of a mutual matrix of unrelated meshes are sums with cancellation: judged against the magnitude of their
    # terms as in test_gpu_coupling.py::test_cross_coupling (the reference's own atomics move them by ~eps * A)
    A = mutual_abs_sum(O1, O2)[np.ix_(O1.pmap[rows] - 1, O2.pmap[cols] - 1)]
    assert (np.abs(Bg - Bo) <= 1e-10 * np.abs(Bo) + 64 * np.finfo(float).eps * A).all()
    well = np.abs(Bo) > 1e-4 * A
    assert (np.abs(Bg - Bo)[well] <= 1e-10 * np.abs(Bo)[well]).all()


def test_lmatblock_strip_ports_scale(env):
    """strip of one DOF against a 6 000-vertex column block of the ports example mesh (22 580 vertices)"""
    O, T = make(env, 'ex_ports')
    rng = np.random.default_rng(9)
    cols = np.sort(rng.permutation(O.np_)[:6000]).astype(np.int32)
    for v in (11, 9000, int(cols[77])):
        close(T.compute_Lmatblock(np.array([v], np.int32), cols), O.lmat_block(np.array([v], np.int32), cols))


@pytest.mark.parametrize('name,js,vc', [('torus', 0, None), ('cyl', 2, None), ('torus', 0, [(1.5, 0.5), (1.5, -0.5)])])
def test_lmathole_vs_oracle(env, name, js, vc):
    O, T = make(env, name, js, vcoils=vc)
    if vc:
        O.compute_Mcoil()
        T.compute_Mcoil()
    Ho = O.lmat_hole()
    Hg = T.compute_LmatHole()
    assert Hg.shape == (O.nholes + O.n_vcoils, O.nelems)
    close(Hg, Ho)


@pytest.mark.parametrize('name,js', [('plate', 0), ('cyl', 2), ('torus', 0)])
def test_bops_block_vs_oracle(env, name, js):
    O, T = make(env, name, js)
    rng = np.random.default_rng(4)
    rows = np.sort(rng.permutation(O.np_)[:120]).astype(np.int32)
    cols = rng.permutation(O.np_)[:300].astype(np.int32)  # includes vertices of the row block's own cells (on-surface)
    Ball = T.compute_Bops_block(rows, cols)  # all three components in one sweep
    Bos = [O.bops_block(rows, cols, d) for d in range(3)]
    sc = max(np.abs(Bo).max() for Bo in Bos)  # one scale for the vector (the flat plate has pure rounding noise in x, y)
    for d in range(3):
        Bo = Bos[d]
        Bd = T.compute_Bops_block(rows, cols, direction=d)
        # same criterion as test_gpu_coupling.py::test_bmat: near entries are central differences of the analytic
        # potential with h = 1e-6, which amplify last-digit differences of log/atan2 by ~1e6 (SURVEY hard part 8)
        assert np.abs(Bd - Bo).max() / sc < 1e-8
        far = np.abs(Bo) < 1e-2 * sc
        assert np.abs(Bd - Bo)[far].max() / sc < 1e-9
        assert np.array_equal(Ball[d], Bd)


def test_cross_eval_vs_oracle(env):
    O1, T1 = make(env, 'plate')
    O2, T2 = make(env, 'cyl', 2)
    rng = np.random.default_rng(21)
    a = rng.standard_normal((5, O1.nelems))  # two groups of right-hand sides
    co, cg = np.zeros(3, np.int64), np.zeros(3, np.int64)
    bo = O1.cross_eval(O2, a, co)
    bg = T1.cross_eval(T2, a, counts=cg)
    assert np.array_equal(co, cg), (co, cg)  # identical quadrature decisions for every cell pair
    # b = M a with random signs in a: entries are sums with cancellation, judged against the largest entry of each field
    # (1e-10) and against themselves at 1e-9
    vclose = lambda x, y: (close(x, y, tol=1e-10, small=10.0), close(x, y, tol=1e-9, small=1e-6))
    vclose(bg, bo)
    vclose(T1.cross_eval(T2, a[:1]), bo[:1])  # through the reference-named entry point
    # a model against itself: coincident cells and shared vertices (all three classes), holes on both sides
    O3, T3 = make(env, 'torus')
    a3 = rng.standard_normal((2, O3.nelems))
    bo3 = O3.cross_eval(O3, a3, co)
    bg3 = T3.cross_eval(T3, a3, counts=cg)
    assert np.array_equal(co, cg) and (co > 0).all(), (co, cg)
    vclose(bg3, bo3)
    with pytest.raises(IndexError):
        T1.cross_eval(T2, a[:, :-1])


def test_reduced_model_file(env, tmp_path):
    """tw_reduce_model: V^T L V, V^T R V, Ms V, V^T Mc, B V from the GPU against numpy on the same operators; the file is
    read back with the HDF5 reader that parses the reference's own fixture meshes."""
    from oracle import h5min
    O, T = make(env, 'torus', icoils=[(1.5, 0.5), (1.5, -0.5)])
    T.compute_Mcoil()
    T.compute_Lmat()
    T.compute_Rmat()
    Ms, Msc, sensor_obj = T.compute_Msensor(sensors=[(ref_floop(R, Z), 1.0) for (R, Z) in [(1.4, 0.0), (0.6, 0.0)]])
    Bmat, Bdr = T.compute_Bmat()
    rng = np.random.default_rng(2)
    V = rng.standard_normal((11, T.nelems))  # > 8 vectors: two passes over the rows
    fn = str(tmp_path / 'reduced.h5')
    T.build_reduced_model(V, filename=fn, compute_B=True, sensor_obj=sensor_obj)
    h = h5min.H5(fn)
    tree = h.tree()
    get = lambda k: np.array(h.read(tree[k]))
    assert int(get('ThinCurr_Version').ravel()[0]) == 1
    assert np.array_equal(get('Basis'), V)
    L = np.asarray(T.Lmat)
    R = T.Rmat.toarray() if hasattr(T.Rmat, 'toarray') else np.asarray(T.Rmat)
    close(get('L'), V @ L @ V.T, tol=1e-11, small=1e-9)
    close(get('R'), V @ R @ V.T, tol=1e-11, small=1e-9)
    Mc = np.asarray(T.compute_Mcoil())  # (n_icoils, nelems)
    close(get('Mc'), Mc @ V.T, tol=1e-11, small=1e-9)
    Bel = np.asarray(Bmat).reshape(3, T.np, T.nelems)  # memory of Fortran Bel(nelems,np,3)
    for k, nm in enumerate(('Bx', 'By', 'Bz')):
        close(get(nm), V @ Bel[k].T, tol=1e-11, small=1e-9)
        assert np.array_equal(get(nm + '_c'), np.asarray(Bdr)[k])
    close(get('Ms'), V @ np.asarray(Ms), tol=1e-11, small=1e-9)
    assert np.array_equal(get('Msc'), np.asarray(Msc))
