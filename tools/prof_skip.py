#!/usr/bin/env python
"""Tuning aid: phase split of lmat_tile_kernel by the skip switches of the TEST build (THINCURR_B200_DEBUG_SKIP:
bit0 near field, bit1 far field, bit2 contraction), all variants in one process.
usage: prof_skip.py <workload> [nshards shard [skips...]]   (library: the -DTW_TEST_HOOKS build)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('THINCURR_B200_LIB', os.path.join(ROOT, 'openfusiontoolkit_b200', 'libthincurr_b200_test.so'))
import torch
import bench
from openfusiontoolkit_b200 import OFT_env
from openfusiontoolkit_b200.ThinCurr import ThinCurr

wl = sys.argv[1] if len(sys.argv) > 1 else 'vessel100k'
nsh = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sh = int(sys.argv[3]) if len(sys.argv) > 3 else 0
skips = [int(s) for s in sys.argv[4:]] or [0, 1, 2, 4, 3, 7]
m = bench.make_mesh(wl)
T = ThinCurr(OFT_env(nthreads=-1))
T.setup_model(r=m['r'], lc=m['lc'], nodesets=m['nodesets'], closures=m['closures'])
out = torch.empty((len(T.shard_rows(nsh, sh)), T.nelems), dtype=torch.float64, device='cuda')
for s in skips:
    os.environ['THINCURR_B200_DEBUG_SKIP'] = str(s)
    best = 1e30
    for r in range(2):
        st = T.compute_Lmat_shard(nsh, sh, out, stream=torch.cuda.current_stream().cuda_stream, stats=True)
        torch.cuda.synchronize()
        best = min(best, (int(st[4]) - int(st[6])) * 1e-6)
    print('%s shard %d/%d skip %d kernel_ms %.2f inv_r %d phipot %d' % (wl, sh, nsh, s, best, st[2], st[3]), flush=True)
