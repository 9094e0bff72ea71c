"""GPU diagnostics (run on the B200 box): device-function accuracy and per-pair parity."""
import sys, os, ctypes, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from helpers import load_mesh, split_nodesets
from oracle import tw_oracle as tw
from openfusiontoolkit_b200 import OFT_env
from openfusiontoolkit_b200 import _interface as I
from openfusiontoolkit_b200.ThinCurr import ThinCurr

rng = np.random.default_rng(0)
x = np.exp(rng.uniform(-30, 10, 200000))
y = np.zeros_like(x)
assert I.b200_probe_rsqrt(len(x), x, y) == 0
print('rsqrt_fast max rel err %.3e' % np.abs(y * np.sqrt(x) - 1).max())

name = sys.argv[1] if len(sys.argv) > 1 else 'plate'
js = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = load_mesh(name)
ns = split_nodesets(m, js)
cl = m['sidesets'][0] if m['sidesets'] else ()
O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=ns, closures=cl)
nc = O.nc
P = O.r[O.lc].reshape(nc, 9)
ii, jj = np.meshgrid(np.arange(nc), np.arange(nc), indexing='ij')
sel = rng.choice(nc * nc, size=min(nc * nc, 400000), replace=False)
ci, cj = ii.ravel()[sel], jj.ravel()[sel]
Pi, Pj = np.ascontiguousarray(P[ci]), np.ascontiguousarray(P[cj])
Ai, Aj = np.ascontiguousarray(O.ca[ci]), np.ascontiguousarray(O.ca[cj])
n = len(ci)
Tg = np.zeros(n); qg = np.zeros(n, np.int32); To = np.zeros(n); qo = np.zeros(n, np.int32)
mode = int(os.environ.get('PROBE_MODE', '1'))
assert I.b200_probe_pairs(n, mode, Pi, Ai, Pj, Aj, Tg, qg) == 0
print('probe mode', mode, 'exact-path fraction %.2e' % ((qg & 64) != 0).mean())
qg = qg & 31
L = tw.lib()
L.tco_pair_T_batch(n, Pi.ctypes.data_as(ctypes.c_void_p), Ai.ctypes.data_as(ctypes.c_void_p), Pj.ctypes.data_as(ctypes.c_void_p),
                   Aj.ctypes.data_as(ctypes.c_void_p), To.ctypes.data_as(ctypes.c_void_p), qo.ctypes.data_as(ctypes.c_void_p))
print('iquad mismatches', (qg != qo).sum(), 'of', n)
rel = np.abs(Tg - To) / np.abs(To)
for q in range(4, 19):
    s = qo == q
    if s.any():
        print('iquad %2d: n=%7d max rel err %.3e' % (q, s.sum(), rel[s].max()))
k = np.argmax(rel)
print('worst pair', ci[k], cj[k], 'iq', qo[k], Tg[k], To[k], rel[k])
np.set_printoptions(precision=17)
print('Pi', Pi[k], 'Pj', Pj[k])

T = ThinCurr(OFT_env(nthreads=-1))
T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=ns, closures=cl if len(cl) else None)
t = time.time(); T.compute_Lmat(); print('gpu L', time.time() - t)
Lo = O.compute_Lmat()
D = np.abs(T.Lmat - Lo)
scale = np.abs(Lo).max()
big = np.abs(Lo) > 1e-8 * scale
R = np.where(big, D / np.maximum(np.abs(Lo), 1e-300), D / scale)
k = np.unravel_index(np.argmax(R), R.shape)
print('L worst entry', k, T.Lmat[k], Lo[k], R[k], 'diag?', k[0] == k[1], 'N', Lo.shape)
print('entries with err>1e-10:', (R > 1e-10).sum(), ' >1e-12:', (R > 1e-12).sum(), 'median', np.median(R))
