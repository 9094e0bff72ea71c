"""world_size-2 gloo test of the multi-rank host logic (bench.py's sharding): every rank plans the
same model, owns a disjoint row block, and the union covers the matrix; the weak-scaling mesh sizes
grow so that pairs per rank stay constant."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from openfusiontoolkit_b200 import OFT_env
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    from openfusiontoolkit_b200.ThinCurr.meshing import build_torus_vessel
    m = build_torus_vessel(40, 80, nports=4)
    T = ThinCurr(OFT_env(nthreads=-1))
    T.setup_model(r=m['r'], lc=m['lc'], nodesets=m['nodesets'], closures=m['closures'])
    rows = T.shard_rows(world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, rows.tolist())
    hashes = [None] * world
    dist.all_gather_object(hashes, T.model_hashes())
    if rank == 0:
        q.put((T.nelems, gathered, hashes))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_partition():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    nelems, gathered, hashes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allr = np.concatenate([np.array(g) for g in gathered])
    assert np.array_equal(np.sort(allr), np.arange(nelems))
    assert len(set(hashes)) == 1, 'all ranks must see the same model'
    assert abs(len(gathered[0]) - len(gathered[1])) < 0.2 * nelems


def _worker_sym(rank, world, port, q):
    """symmetric partition + exchange on CPU tensors (gloo): every rank holds the entries of a known symmetric matrix
    that the symmetric build evaluates on it (diagonal block + its checkerboard half of the shared blocks) and must end
    up with its complete rows."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from openfusiontoolkit_b200 import OFT_env
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    from openfusiontoolkit_b200.ThinCurr.meshing import build_torus_vessel
    m = build_torus_vessel(40, 80, nports=4)
    T = ThinCurr(OFT_env(nthreads=-1))
    T.setup_model(r=m['r'], lc=m['lc'], nodesets=m['nodesets'], closures=m['closures'])
    N = T.nelems
    ids = [T.shard_rows_sym(world, s) for s in range(world)]
    rng = np.random.default_rng(7)
    A = rng.standard_normal((N, N))
    A = A + A.T
    mine = ids[rank]
    mask = T.sym_computed_mask(world, rank, ids)
    out = torch.from_numpy(np.where(mask, A[mine], 0.0))   # what the symmetric build leaves for the exchange is zero
    T.exchange_symmetric(out, world, rank, row_ids=ids)
    ok = bool(np.array_equal(out.numpy(), A[mine])) and 0.3 < mask.mean() < 0.8
    res = [None] * world
    dist.all_gather_object(res, (ok, [len(i) for i in ids], int(sum(len(i) for i in ids)), N))
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_symmetric_exchange(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() + 17 * world) % 2000
    procs = [ctx.Process(target=_worker_sym, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for ok, sizes, total, N in res:
        assert ok, 'rows incomplete after the exchange'
        assert total == N
        assert max(sizes) < 1.6 * min(sizes), 'shards own similar numbers of rows'


def test_bench_workloads_are_the_baseline_configs():
    """bench.py's meshes: configs[3] (~100k vertices, the default at every N: strong scaling), configs[1] (~20k),
    configs[4] (~150k, ~180 GB matrix)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.WORKLOADS['vessel100k'] == (224, 448)
    m = bench.make_mesh('vessel20k')
    assert 19000 < m['r'].shape[0] < 21000
    nt, nphi = bench.WORKLOADS['vessel150k']
    assert 170e9 < (nt * nphi) ** 2 * 8 < 190e9
