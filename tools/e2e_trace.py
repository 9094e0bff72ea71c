"""Trace of the reference-facing build (thincurr_Lmat) over NDEV devices of one process:
usage e2e_trace.py <workload> <ndev> [patch sizes...]   (patch sizes: tuning of the streamed single-device build)"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['THINCURR_B200_TRACE'] = '1'
os.environ['THINCURR_B200_NDEV'] = sys.argv[2] if len(sys.argv) > 2 else '1'
import bench
from openfusiontoolkit_b200 import OFT_env
from openfusiontoolkit_b200.ThinCurr import ThinCurr
m = bench.make_mesh(sys.argv[1] if len(sys.argv) > 1 else 'vessel100k')
T = ThinCurr(OFT_env(nthreads=-1))
T.setup_model(r=m['r'], lc=m['lc'], nodesets=m['nodesets'], closures=m['closures'])
for P in (sys.argv[3:] or [None]):
    if P is not None:
        os.environ['THINCURR_B200_PATCH'] = P
    for rep in range(3 if P is None else 2):
        t = time.perf_counter()
        T.compute_Lmat()
        print('compute_Lmat (patch %s) call %d: %.3f s' % (P, rep, time.perf_counter() - t), flush=True)
