"""Row-block sharding of the dense L build and full-size checks on the benchmark
meshes: shards partition the rows, a sharded build reproduces the single-device matrix bit for bit
(owner-computes tiles => deterministic), the matrix is exactly symmetric, and random rows of the
large matrices match the oracle's per-entry definition to <= 1e-10."""
import numpy as np
import pytest
from helpers import load_mesh, split_nodesets
from oracle import tw_oracle as tw

pytestmark = pytest.mark.gpu
ENTRY_TOL = 1e-10


@pytest.fixture(scope='module')
def env():
    from openfusiontoolkit_b200 import OFT_env
    return OFT_env(nthreads=-1)


def entry_err(A, B, scale):
    big = np.abs(B) > 1e-8 * scale
    rel = (np.abs(A - B)[big] / np.abs(B)[big]).max() if big.any() else 0.0
    ab = (np.abs(A - B)[~big] / scale).max() if (~big).any() else 0.0
    return max(rel, ab)


@pytest.mark.parametrize('nshards', [2, 3, 8])
def test_sharded_equals_full(env, nshards):
    import torch
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('ex_torus')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0] if m['sidesets'] else None)
    T.compute_Lmat()
    full = np.array(T.Lmat)
    assert np.array_equal(full, full.T)
    seen = np.zeros(T.nelems, bool)
    for s in range(nshards):
        rows = T.shard_rows(nshards, s)
        assert not seen[rows].any()
        seen[rows] = True
        out = torch.empty((len(rows), T.nelems), dtype=torch.float64, device='cuda')
        T.compute_Lmat_shard(nshards, s, out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), full[rows]), 'shard %d differs from the single-device build' % s
        host = np.zeros((len(rows), T.nelems))
        T.compute_Lmat_shard_host(nshards, s, host)
        assert np.array_equal(host, full[rows])
    assert seen.all()


@pytest.mark.parametrize('nshards', [2, 5])
def test_symmetric_shards_equal_full(env, nshards):
    """symmetric shards (thincurr_b200_Lmat_shard_sym): the row sets partition the DOFs; every entry is evaluated on
    exactly one shard (its diagonal block, and a checkerboard half of every block shared with another shard); what a
    shard evaluates is the single-device matrix -- bit for bit where the tile keeps the single-device orientation (row
    patch < column patch), to rounding of the summation order otherwise -- and the rest is left zero for the exchange."""
    import torch
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('ex_torus')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0] if m['sidesets'] else None)
    T.compute_Lmat()
    full = np.array(T.Lmat)
    scale = np.abs(full).max()
    ids = [T.shard_rows_sym(nshards, s) for s in range(nshards)]
    assert np.array_equal(np.sort(np.concatenate(ids)), np.arange(T.nelems))
    covered = np.zeros((T.nelems, T.nelems), np.int32)
    sizes = []
    for s in range(nshards):
        out = torch.empty((len(ids[s]), T.nelems), dtype=torch.float64, device='cuda')
        T.compute_Lmat_shard_sym(nshards, s, out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        mask = T.sym_computed_mask(nshards, s, ids)
        ref = full[ids[s]]
        assert not got[~mask].any(), 'shard %d wrote entries that belong to another shard' % s
        assert np.abs(got[mask] - ref[mask]).max() <= 1e-13 * scale, 'shard %d' % s
        own = np.isin(np.arange(T.nelems), ids[s])
        assert np.array_equal(got[:, own], ref[:, own]), 'diagonal block of shard %d must be the single-device bits' % s
        covered[ids[s]] += mask
        sizes.append(mask.sum())
    # every entry evaluated exactly once: an entry and its transpose never both on different shards
    owner = np.zeros(T.nelems, np.int32)
    for s, i in enumerate(ids):
        owner[i] = s
    off = owner[:, None] != owner[None, :]
    assert np.array_equal((covered + covered.T)[off], np.ones(off.sum(), np.int32)), 'off-diagonal-block entries: exactly one of (a,b), (b,a)'
    assert covered[~off].all()
    assert max(sizes) / (sum(sizes) / nshards) < 1.25, 'shards evaluate similar numbers of entries'


def test_ports_mesh_rows(env):
    """BASELINE config 2 (ports mesh, 22 580 vertices / 44 560 triangles, 11 holes): full dense L on
    one GPU; 40 random vertex rows and all hole rows against the oracle."""
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('ex_ports')
    ns = m['nodesets']
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=ns)
    T.compute_Lmat()
    L = T.Lmat
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=ns)
    assert O.nelems == T.nelems
    rng = np.random.default_rng(5)
    rows = np.concatenate([rng.choice(O.np_active, 40, replace=False), np.arange(O.np_active, O.nelems)]).astype(np.int32)
    Ro = O.lmat_rows(rows)
    scale = np.abs(np.diag(L)).max()
    err = entry_err(np.array(L[rows]), Ro, scale)
    assert err < ENTRY_TOL, 'max rel entry error %.3e' % err
    # symmetry without materialising a transpose copy of the 3.9 GB matrix
    for r in rows[:20]:
        assert np.array_equal(L[r, :], L[:, r])


def test_vessel_mesh_rows_and_stats(env):
    """The benchmark workload (synthetic vessel, ~20k vertices): row parity, symmetry, and the
    device's pair statistics against the oracle's loop-nest counts on a row-cell sample."""
    from bench import make_mesh
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = make_mesh('vessel20k')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], nodesets=m['nodesets'], closures=m['closures'])
    T.compute_Lmat()
    L = T.Lmat
    O = tw.OracleModel(m['r'], m['lc'], None, nodesets=m['nodesets'], closures=m['closures'])
    rng = np.random.default_rng(9)
    rows = np.concatenate([rng.choice(O.np_active, 24, replace=False), np.arange(O.np_active, O.nelems)]).astype(np.int32)
    Ro, A = O.lmat_rows(rows, with_abs=True)
    got = np.array(L[rows])
    # criterion of tests/test_gpu_scale.py: 1e-10 relative on every entry plus the summation-order noise of an entry that
    # cancels to a small fraction of its terms (A = sum of their magnitudes; the reference's atomics reorder them from run
    # to run); plain 1e-10 where fewer than 3 digits are lost
    assert (np.abs(got - Ro) <= 1e-10 * np.abs(Ro) + 64 * np.finfo(float).eps * A).all()
    well = np.abs(Ro) > 1e-3 * A
    assert (np.abs(got - Ro)[well] / np.abs(Ro)[well]).max() < ENTRY_TOL
    for r in rows[:12]:
        assert np.array_equal(L[r, :], L[:, r])
    hist, visited = T.pair_stats()
    assert visited == hist.sum() and hist[:4].sum() == 0 and 0 < visited <= O.nc ** 2


@pytest.mark.parametrize('name,js', [('plate', 0), ('cyl', 2), ('torus', 0)])
def test_pair_stats_match_reference_loop(env, name, js):
    """Visited-pair count and order histogram of the reference loop nest (thin_wall.F90:1028-1059),
    the denominator of the benchmark metric: GPU count == oracle count, bin by bin."""
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh(name)
    ns = split_nodesets(m, js)
    cl = m['sidesets'][0] if m['sidesets'] else None
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=ns if len(ns) else None, closures=cl)
    hist, visited = T.pair_stats()
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=ns, closures=cl if cl is not None else ())
    h2 = np.zeros(19, np.int64)
    O.compute_Lmat(hist=h2)
    assert O.visited == visited
    assert np.array_equal(h2, hist)


def test_dense_block_evaluator(env):
    """thincurr_b200_Lmat_block: L restricted to arbitrary row / column DOF subsets (the dense-block evaluator of a
    hierarchical compression, thin_wall_hodlr.F90:136-404) against the full matrix: same pair integrals and roles, only
    the summation order of an entry may differ from the mirrored full build (<= 1e-13 of the entry scale)."""
    import torch
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('ex_torus')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0] if m['sidesets'] else None)
    T.compute_Lmat()
    full = np.array(T.Lmat)
    N = T.nelems
    rng = np.random.default_rng(11)
    cases = [(rng.choice(N, 300, replace=False), rng.choice(N, 500, replace=False)),      # scattered subsets
             (np.arange(100, 420), np.arange(100, 420)),                                   # a diagonal block
             (np.array([N - 1, 0, N - 2]), np.arange(N)[::-1].copy()),                     # hole rows, all columns reversed
             (np.arange(N), np.array([5]))]                                                # one column
    scale = np.abs(full).max()
    for rows, cols in cases:
        out = torch.full((len(rows), len(cols) + 3), -7.0, dtype=torch.float64, device='cuda')   # ld > ncols: padding untouched
        T.compute_Lmat_block(rows, cols, out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert np.all(got[:, len(cols):] == -7.0)
        ref = full[np.ix_(rows, cols)]
        assert np.abs(got[:, :len(cols)] - ref).max() <= 1e-13 * scale
        big = np.abs(ref) > 1e-8 * scale
        assert (np.abs(got[:, :len(cols)] - ref)[big] / np.abs(ref)[big]).max() < 1e-10
    with pytest.raises(Exception):
        T.compute_Lmat_block([0, 0], [1], torch.empty((2, 1), dtype=torch.float64, device='cuda'))


@pytest.mark.parametrize('name,bands', [('torus', '3'), ('ex_torus', '6')])
def test_streamed_single_device_build(env, monkeypatch, name, bands):
    """The reference-facing build of a large model on one device (thincurr_Lmat -> lmat_stream_host): banded plan over ranges
    of reference ids, the matrix in the reference layout on the device, every band followed by its mirror pass and two
    strided device->host copies.  Forced onto small meshes here (holes, closure, periodic torus): entries vs the oracle,
    exact symmetry, the same bits from the device-resident shard builds of the same plan and from the pageable-destination
    path (row bands over the same patches)."""
    import torch
    from openfusiontoolkit_b200 import _interface as I
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    monkeypatch.setenv('THINCURR_B200_STREAM_BANDS', bands)
    m = load_mesh(name)
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0])
    T.compute_Lmat()
    L = np.array(T.Lmat)
    assert np.array_equal(L, L.T)
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0])
    Lo = O.compute_Lmat()
    err = entry_err(L, Lo, np.abs(np.diag(Lo)).max())
    assert err < ENTRY_TOL, 'max rel entry error %.3e' % err
    for s in range(2):
        rows = T.shard_rows(2, s)
        out = torch.empty((len(rows), T.nelems), dtype=torch.float64, device='cuda')
        T.compute_Lmat_shard(2, s, out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), L[rows]), 'shard %d differs from the streamed build' % s
    P = np.full((T.nelems, T.nelems), np.nan)   # pageable caller memory
    assert I.b200_Lmat_host(T.tw_obj, P) == 0, I.b200_last_error()
    assert np.array_equal(P, L)
