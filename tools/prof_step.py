#!/usr/bin/env python
"""Tuning aid: build the L matrix of a bench workload once or a few times and print the tile kernel's duration
(device globaltimer) and the device-side evaluation counters.  usage: prof_step.py <workload> [reps [nshards shard]]
Library variant through THINCURR_B200_LIB (e.g. the -DTW_LMAT_PROF or -DTW_TEST_HOOKS builds)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from openfusiontoolkit_b200 import OFT_env
from openfusiontoolkit_b200.ThinCurr import ThinCurr

wl = sys.argv[1] if len(sys.argv) > 1 else 'vessel100k'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nsh = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # build only shard `sh` of `nsh` (full rows): short kernels for ncu
sh = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m = bench.make_mesh(wl)
T = ThinCurr(OFT_env(nthreads=-1))
T.setup_model(r=m['r'], lc=m['lc'], nodesets=m['nodesets'], closures=m['closures'])
N = T.nelems
out = torch.empty((len(T.shard_rows(nsh, sh)), N), dtype=torch.float64, device='cuda')
for r in range(reps):
    st = T.compute_Lmat_shard(nsh, sh, out, stream=torch.cuda.current_stream().cuda_stream, stats=True)
    torch.cuda.synchronize()
    print('%s %s kernel_ms %.2f far_pairs %d near_T %d inv_r %d phipot %d' % (wl, os.environ.get('THINCURR_B200_DEBUG_SKIP', '-'), (int(st[4]) - int(st[6])) * 1e-6, st[0], st[1], st[2], st[3]), flush=True)
print('checksum %.10e' % float(out.sum()))
