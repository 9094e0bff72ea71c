"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run): symmetric shards built on
every rank, the library's exchange over cudaIpc-mapped peer memory (thincurr_b200_Lmat_exchange), the gather of the full
matrix on rank 0 (thincurr_b200_Lmat_gather) and a sharded mat-vec / Lanczos eigen solve on the resident row blocks;
everything is compared with the single-device build on rank 0."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.distributed as dist
from helpers import MU0, load_mesh
from openfusiontoolkit_b200 import OFT_env
from openfusiontoolkit_b200.ThinCurr import ThinCurr


class DevBuf:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 3, 'strides': None}


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    m = load_mesh(sys.argv[1] if len(sys.argv) > 1 else 'ex_torus')
    T = ThinCurr(OFT_env(nthreads=-1))
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0] if m['sidesets'] else None)
    T.set_eta_values(eta_surf=np.array([10.0 * MU0]))
    N = T.nelems
    ids = [T.shard_rows_sym(world, s) for s in range(world)]
    nrows = len(ids[rank])
    ptr = T.device_alloc(max(nrows, 1) * N * 8)
    out = torch.as_tensor(DevBuf(ptr, (nrows, N)), device='cuda')
    handles = [None] * world
    dist.all_gather_object(handles, T.ipc_export(ptr))
    peers = [ptr if s == rank else T.ipc_open(handles[s]) for s in range(world)]
    stream = torch.cuda.current_stream().cuda_stream
    token = torch.zeros(1, device='cuda')
    for rep in range(2):  # twice: the second build overwrites rows the peers have read
        T.compute_Lmat_shard_sym(world, rank, out, stream=stream)
        dist.all_reduce(token)
        T.exchange_symmetric_peer(ptr, N, world, rank, peers, stream=stream)
        dist.all_reduce(token)
    torch.cuda.synchronize()
    ok = True
    msgs = []
    # sharded mat-vec + Lanczos on the resident rows (every rank calls; y is all-gathered)
    T.compute_Rmat()
    xd = torch.empty(N, dtype=torch.float64, device='cuda')
    yd = torch.empty(max(nrows, 1), dtype=torch.float64, device='cuda')
    order = np.concatenate(ids)

    def apply(x):
        xd.copy_(torch.from_numpy(np.ascontiguousarray(x)))
        T.rows_apply(ptr, N, nrows, N, xd.data_ptr(), yd.data_ptr(), stream=stream)
        parts = [torch.empty(len(i), dtype=torch.float64, device='cuda') for i in ids]
        dist.all_gather(parts, yd[:nrows])
        y = np.empty(N)
        y[order] = torch.cat(parts).cpu().numpy()
        return y
    vals, vecs, napp = T.get_eigs_sharded(4, apply)
    if rank == 0:
        full_ptr = T.device_alloc(N * N * 8)
        T.gather_full(world, True, peers, N, full_ptr, N, stream=stream)
        torch.cuda.synchronize()
        full = torch.as_tensor(DevBuf(full_ptr, (N, N)), device='cuda').cpu().numpy()
        os.environ['THINCURR_B200_NDEV'] = '1'
        T.compute_Lmat()
        ref = np.array(T.Lmat)
        # a tile evaluated with the larger patch index as row side sums in another order than the single-device build:
        # equal to rounding, and exactly symmetric (every entry is evaluated once and mirrored)
        if not (np.abs(full - ref).max() <= 1e-13 * np.abs(ref).max()):
            ok = False
            msgs.append('gathered matrix differs from the single-device build: max abs %.3e' % np.abs(full - ref).max())
        if not np.array_equal(full, full.T):
            ok = False
            msgs.append('gathered matrix is not exactly symmetric')
        if N <= 4000:   # dense generalised eigen solve on the host (torchrun pins OMP_NUM_THREADS=1: small meshes only)
            import scipy.linalg as sl
            w = np.sort(sl.eigh(ref, T.Rmat.toarray(), eigvals_only=True))[::-1][:4]
            if np.abs(vals / w - 1.0).max() > 1e-8:
                ok = False
                msgs.append('sharded Lanczos eigenvalues off: %s vs %s' % (vals, w))
        for k in range(4):   # eigen residuals against the single-device matrix
            Lv = ref @ vecs[k]
            res = np.linalg.norm(Lv - vals[k] * (T.Rmat @ vecs[k])) / np.linalg.norm(Lv)
            if res > 1e-7:
                ok = False
                msgs.append('eigenpair %d residual %.2e' % (k, res))
        if not (np.all(np.diff(vals) <= 0) and vals[0] > 0):
            ok = False
            msgs.append('eigenvalues not the leading ones in descending order: %s' % vals)
        msgs.append('eigs %s in %d mat-vecs' % (vals, napp))
        T.device_free(full_ptr)
    flag = torch.tensor([1.0 if ok else 0.0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    for s in range(world):
        if s != rank:
            T.ipc_close(peers[s])
    dist.barrier()
    del out
    T.device_free(ptr)
    if rank == 0:
        print('\n'.join(msgs))
        print('MULTI_OK' if flag.item() == 1.0 else 'MULTI_FAIL')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
