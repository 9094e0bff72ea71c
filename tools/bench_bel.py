"""Time the B-field operator build (Bel rows, tw_compute_Bops) for the benchmark vessel on cuda:0.
usage: python tools/bench_bel.py [nshards]   (times shard 0 of nshards; default 1 = all elements)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_mesh
from openfusiontoolkit_b200 import OFT_env
from openfusiontoolkit_b200.ThinCurr import ThinCurr
nsh = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mesh = make_mesh(1, 'auto')
T = ThinCurr(OFT_env(nthreads=-1))
T.setup_model(r=mesh['r'], lc=mesh['lc'], nodesets=mesh['nodesets'], closures=mesh['closures'])
rows = T.shard_rows(nsh, 0)
npts = mesh['r'].shape[0]
out = torch.empty((3, npts, len(rows)), dtype=torch.float64, device='cuda')
stream = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    T.compute_Bel_shard(nsh, 0, out, stream=stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    T.compute_Bel_shard(nsh, 0, out, stream=stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
nc = mesh['lc'].shape[0]
print('Bel shard 0/%d: %d elements x %d vertices, %.1f ms, %.3e (cell,vertex) pairs/s, output %.2f GB' % (
    nsh, len(rows), npts, ms, nc * npts / nsh / (ms * 1e-3), out.numel() * 8 / 1e9))
