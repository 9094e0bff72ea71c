"""B-field operator build (Bel rows, tw_compute_Bops, thin_wall.F90:1989-2112) on cuda:0: one JSON line in bench.py's
format (device-timed value, roofline against the builder-measured FP64 peak, CPU baseline from the oracle's loop).
usage: python tools/bench_bel.py [workload [nshards]]   (times shard 0 of nshards; default vessel20k, 1 = all elements)

Unit of work: one (cell, vertex) pair = the field of one triangle's three basis currents at one mesh vertex.
Algorithmic flops per pair (SURVEY 8d conventions: add/sub/mul = 1, FMA = 2, sqrt = div = 1):
  order selection 3 x (3 sub + 3 mul + 2 add + sqrt) + 5 = 32; far pair with an n-point rule: per point
  9 (point) + 3 (difference) + 5 (|r|^2) + 3 (r^-3: sqrt, mul, div) + 3 x (6 (cross) + 3 (scale) + 3 (sum)) = 56,
  i.e. F_far(n) = 56 n + 9; near pair (finite differences of the analytic potential): 6 (or 12 on-surface) potential
  evaluations of ~230 flops; scatter 3 x 3 = 9.  The rule histogram of the mesh is measured by the oracle's
  selection on a row sample."""
import ctypes
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from openfusiontoolkit_b200 import OFT_env
from openfusiontoolkit_b200 import _interface as I
from openfusiontoolkit_b200.ThinCurr import ThinCurr

wl = sys.argv[1] if len(sys.argv) > 1 else 'vessel20k'
nsh = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mesh = bench.make_mesh(wl)
T = ThinCurr(OFT_env(nthreads=-1))
T.setup_model(r=mesh['r'], lc=mesh['lc'], nodesets=mesh['nodesets'], closures=mesh['closures'])
rows = T.shard_rows(nsh, 0)
npts, nc = mesh['r'].shape[0], mesh['lc'].shape[0]
out = torch.empty((3, npts, len(rows)), dtype=torch.float64, device='cuda')
stream = torch.cuda.current_stream().cuda_stream
l0 = I.b200_launch_count()
for _ in range(3):
    T.compute_Bel_shard(nsh, 0, out, stream=stream)
torch.cuda.synchronize()
launches = (I.b200_launch_count() - l0) // 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n):
    T.compute_Bel_shard(nsh, 0, out, stream=stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
pairs = nc * npts / nsh

# order histogram + CPU baseline: the oracle's tw_compute_Bops element loop on a slab of cells
from oracle import tw_oracle as tw
O = tw.OracleModel(mesh['r'], mesh['lc'], None, nodesets=mesh['nodesets'], closures=mesh['closures'])
ncpu = min(nc, 1600)
i0 = (nc - ncpu) // 2
t0 = time.perf_counter()
O.compute_Bmat(i0, i0 + ncpu, finalize=False)
cpu_dt = time.perf_counter() - t0
cpu_rate = ncpu * npts / cpu_dt
# histogram of the selection rule (thin_wall.F90:2034-2047) on the same slab, in numpy
r, lc = np.asarray(mesh['r'], float), O.lc
qnp = {4: 6, 5: 7, 6: 12, 7: 15, 8: 16, 9: 19, 10: 25}
flops = 0.0
cnt_far = cnt_near = 0
for c in range(i0, i0 + ncpu, 8):
    P = r[lc[c]]
    d = np.linalg.norm(P[:, None, :] - r[None, :, :], axis=2)
    dmin, dmax = d.min(0), np.maximum(d.max(0), np.sqrt(np.maximum(O.ca[c], O.va / np.pi ** 2)))
    with np.errstate(divide='ignore', invalid='ignore'):
        iq = np.where(dmin < 1e-8, 18, np.clip(np.abs(np.trunc(np.log(1e-8) / np.log(1.0 - dmin / dmax))), 4, 18)).astype(int)
    for q, npq in qnp.items():
        k = int((iq == q).sum())
        flops += k * (32 + 56 * npq + 9 + 9)
        cnt_far += k
    kn = int((iq > 10).sum())
    kon = int((dmin < 1e-8).sum())
    flops += (kn - kon) * (32 + 6 * 230 + 30 + 9) + kon * (32 + 12 * 230 + 40 + 9)
    cnt_near += kn
flops_per_pair = flops / (cnt_far + cnt_near)
peak_clock = np.zeros(1)
peak_tf = float(I.b200_dfma_peak(0, peak_clock.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
achieved = flops_per_pair * pairs / (ms * 1e-3) / 1e12
line = {'metric': 'Bel (cell,vertex) pair-fields/s', 'value': pairs / (ms * 1e-3), 'unit': 'pairs/s', 'n_gpus': 1, 'steps': n, 'warmup': 3,
        'ms_per_step': ms, 'higher_is_better': True, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': bench.workload_name(mesh).replace('self-inductance L', 'B-field operator Bel'), 'shard': '0 of %d' % nsh,
                   'elements': int(len(rows)), 'vertices': int(npts), 'output_GB': out.numel() * 8 / 1e9,
                   'near_fraction': cnt_near / (cnt_far + cnt_near)},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                     'flops_per_pair': flops_per_pair, 'kernel': 'bel_tile_kernel',
                     'peak_source': 'builder-measured: DFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 figure)',
                     'hbm_write_GBps': out.numel() * 8 / (ms * 1e-3) / 1e9},
        'cpu_baseline': {'value': cpu_rate, 'unit': 'pairs/s', 'cores': int(tw.lib().tco_num_threads()), 'kind': 'port',
                         'sample': 'cells [%d,%d) x all %d vertices in %.1f s, gcc -O2 -fopenmp schedule(dynamic,100)' % (i0, i0 + ncpu, npts, cpu_dt)}}
print(json.dumps(line), flush=True)
