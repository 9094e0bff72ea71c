"""Timings of the SURVEY 8f rows (csrc/tw_blocks.cu) on cuda:0 next to the oracle's loops on the host cores: one JSON
line per operation (wall clock around the synchronous call, host output included -- this is what an ACA+ / HODLR host
code sees).  usage: python tools/bench_blocks.py [workload]   (default vessel20k)

 - strip: tw_compute_Lmatblock for ONE row DOF against a 2 000-vertex column block (the ACA+ access pattern,
   thin_wall_hodlr.F90:1260-1283), mean over 50 calls
 - block: a 1 500 x 1 500 vertex near-field block (host and device output)
 - hole : tw_compute_LmatHole (all hole columns)
 - bops : tw_compute_Bops_block, 1 500 x 1 500, all three components in one sweep
 - mf   : tw_compute_Lmat_MF (ThinCurr.cross_eval) from the plate test mesh (902 cells, translated into the vessel) onto
          the vessel, 4 right-hand sides -- the plasma-mode -> wall use of the reference
Unit: cell pairs (cell x vertex pairs for bops) per second."""
import ctypes
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import bench
from helpers import load_mesh
from openfusiontoolkit_b200 import OFT_env
from openfusiontoolkit_b200.ThinCurr import ThinCurr
from oracle import tw_oracle as tw

wl = sys.argv[1] if len(sys.argv) > 1 else 'vessel20k'
mesh = bench.make_mesh(wl)
env = OFT_env(nthreads=-1)
T = ThinCurr(env)
T.setup_model(r=mesh['r'], lc=mesh['lc'], nodesets=mesh['nodesets'], closures=mesh['closures'])
O = tw.OracleModel(mesh['r'], mesh['lc'], None, nodesets=mesh['nodesets'], closures=mesh['closures'])
rng = np.random.default_rng(1)
act = np.nonzero(O.pmap > 0)[0].astype(np.int32)
order = act[np.argsort(np.arctan2(O.r[act, 1], O.r[act, 0]))]  # spatially compact blocks: sorted by toroidal angle
A, B = np.sort(order[:1500]), np.sort(order[1500:3000])
cols2k = np.sort(order[4000:6000])
nthreads = int(tw.lib().tco_num_threads())


def timed(fn, n=1):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def cpu(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def line(op, pairs, gpu_s, cpu_s, cpu_pairs, note):
    print(json.dumps({'op': op, 'workload': wl, 'pairs': int(pairs), 'gpu_ms': round(gpu_s * 1e3, 4), 'gpu_pairs_per_s': pairs / gpu_s,
                      'cpu_pairs_per_s': cpu_pairs / cpu_s, 'cpu_threads': nthreads, 'ratio': (pairs / gpu_s) / (cpu_pairs / cpu_s),
                      'note': note}), flush=True)


ncell = lambda pts: len(O.block(pts)[0])
row1 = A[700:701]
p = ncell(row1) * ncell(cols2k)
line('Lmatblock strip 1 x 2000', p, timed(lambda: T.compute_Lmatblock(row1, cols2k), 50), cpu(lambda: O.lmat_block(row1, cols2k)), p,
     'host output, per call')
c = cpu(lambda: O.lmat_block(A[:100], B))
cp = ncell(A[:100]) * ncell(B)
line('Lmatblock 1500 x 1500', ncell(A) * ncell(B), timed(lambda: T.compute_Lmatblock(A, B), 3), c, cp, 'CPU: 100-row sample')
d = torch.zeros((1500, 1500), dtype=torch.float64, device='cuda')
line('Lmatblock 1500 x 1500 (device output)', ncell(A) * ncell(B), timed(lambda: T.compute_Lmatblock(A, B, out=d), 3), c, cp,
     'CPU: 100-row sample')
# hole columns: CPU = the pair integrals of 64 hole cells against all cells (tco_pair_T, the inner loop of tw_compute_LmatHole)
hc = np.nonzero(np.diff(O.kfh) > 0)[0]
sub = hc[:64]
P = np.ascontiguousarray(O.r[O.lc].reshape(O.nc, 9))
ii, jj = np.meshgrid(sub, np.arange(O.nc), indexing='ij')
Pa, Pb = np.ascontiguousarray(P[ii.ravel()]), np.ascontiguousarray(P[jj.ravel()])
Aa, Ab = np.ascontiguousarray(O.ca[ii.ravel()]), np.ascontiguousarray(O.ca[jj.ravel()])
To, qo = np.zeros(len(Aa)), np.zeros(len(Aa), np.int32)
vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
c = cpu(lambda: tw.lib().tco_pair_T_batch(len(Aa), vp(Pa), vp(Aa), vp(Pb), vp(Ab), vp(To), vp(qo)))
line('LmatHole (%d holes, %d hole cells)' % (O.nholes, len(hc)), len(hc) * O.nc, timed(lambda: T.compute_LmatHole(), 1), c, len(Aa),
     'CPU: pair integrals of 64 hole cells x all cells')
line('Bops_block 1500 x 1500 x 3', ncell(A) * len(B), timed(lambda: T.compute_Bops_block(A, B), 3), cpu(lambda: O.bops_block(A[:100], B, 0)),
     ncell(A[:100]) * len(B), 'CPU: 100-row sample, ONE component (the reference makes three calls)')
# matrix-free apply: plate (scaled into the vessel's bore) -> vessel
pm = load_mesh('plate')
rp = pm['r'] * 0.3 + np.array([1.0, 0.0, 0.0])
Tp = ThinCurr(env)
Tp.setup_model(r=rp, lc=pm['lc'], reg=pm['reg'])
Op = tw.OracleModel(rp, pm['lc'], pm['reg'])
a = rng.standard_normal((4, Tp.nelems))
cnt = np.zeros(3, np.int64)
g = timed(lambda: Tp.cross_eval(T, a, counts=cnt), 2)
c = cpu(lambda: Op.cross_eval(O, a))
line('cross_eval plate -> vessel, 4 rhs (classes far/close/vclose = %s)' % cnt.tolist(), Op.nc * O.nc, g, c, Op.nc * O.nc, 'CPU: the whole apply')
