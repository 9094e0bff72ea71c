"""GPU parity: CUDA operator builds (through the C ABI / ThinCurr host class) vs the CPU oracle.

Tolerances: the north star asks for max relative error <= 1e-10 on L entries and <= 1e-8 on the
leading L/R eigenvalues.  Entry errors are measured relative to the entry itself for all entries
larger than 1e-8 * max|L| (smaller ones are sums that cancel to rounding) and relative to max|L|
for the rest.
"""
import os
import numpy as np
import pytest
from helpers import MU0, goldens, load_mesh, split_nodesets, ref_circle, ref_floop, dummy_mesh
from oracle import tw_oracle as tw

pytestmark = pytest.mark.gpu
G = goldens()
ENTRY_TOL = 1e-10
EIG_TOL = 1e-8


def entry_err(A, B):
    scale = np.abs(B).max()
    big = np.abs(B) > 1e-8 * scale
    rel = np.abs(A - B)[big] / np.abs(B)[big]
    absr = np.abs(A - B)[~big] / scale if (~big).any() else np.zeros(1)
    return max(rel.max(), absr.max())


@pytest.fixture(scope='module')
def env():
    from openfusiontoolkit_b200 import OFT_env
    return OFT_env(nthreads=-1)


def build_pair(env, name, jumper_start=0, eta=10.0):
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh(name)
    ns = split_nodesets(m, jumper_start)
    cl = m['sidesets'][0] if m['sidesets'] else None
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=ns, closures=cl if cl is not None else (), eta=[eta * MU0])
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=ns, closures=cl)
    T.set_eta_values(eta_surf=np.array([eta * MU0]))
    return O, T


@pytest.mark.parametrize('name,js', [('plate', 0), ('cyl', 2), ('torus', 0), ('ex_torus', 0), ('ex_cyl', -1)])
def test_lmat_entries_and_eigs(env, name, js):
    import scipy.linalg as sl
    O, T = build_pair(env, name, js)
    Lo = O.compute_Lmat()
    T.compute_Lmat()
    Lg = T.Lmat
    assert Lg.shape == Lo.shape
    assert np.array_equal(Lg, Lg.T), 'L must be exactly symmetric (mirrored like the reference)'
    err = entry_err(Lg, Lo)
    assert err < ENTRY_TOL, 'max rel entry error %.3e' % err
    T.compute_Rmat()
    Ro = O.compute_Rmat().toarray()
    Rg = T.Rmat.toarray()
    assert np.abs(Rg - Ro).max() <= 1e-13 * np.abs(Ro).max()
    if name in ('plate', 'cyl', 'torus', 'ex_cyl'):   # ex_cyl = BASELINE config 0 (src/examples/ThinCurr/cyl)
        wg = np.sort(sl.eigh(Lg, Rg, eigvals_only=True))[::-1][:4]
        wo = np.sort(sl.eigh(Lo, Ro, eigvals_only=True))[::-1][:4]
        assert np.abs(wg / wo - 1.0).max() < EIG_TOL
    if name in ('plate', 'cyl', 'torus'):
        g = G['eig_' + name]
        assert np.abs(wg / np.array(g['vals']) - 1.0).max() < g['tol']


_DRAIN_SCRIPT = r"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'tests'))
from helpers import load_mesh
from openfusiontoolkit_b200 import OFT_env
from openfusiontoolkit_b200.ThinCurr import ThinCurr
m = load_mesh('torus')
T = ThinCurr(OFT_env(nthreads=-1))
T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0])
T.compute_Lmat()
ref = np.array(T.Lmat)
rows = T.shard_rows(3, 2)
out0 = torch.empty((len(rows), T.nelems), dtype=torch.float64, device='cuda')
T.compute_Lmat_shard(3, 2, out0)
torch.cuda.synchronize()
os.environ['THINCURR_B200_DRAIN_LIMIT'] = sys.argv[2]
full = torch.empty((T.nelems, T.nelems), dtype=torch.float64, device='cuda')
T.compute_Lmat_shard(1, 0, full)
out1 = torch.empty_like(out0)
T.compute_Lmat_shard(3, 2, out1)
torch.cuda.synchronize()
allrows = T.shard_rows(1, 0)
assert np.array_equal(full.cpu().numpy(), ref[allrows])
assert np.array_equal(out1.cpu().numpy(), out0.cpu().numpy())
print('DRAIN_OK')
"""


@pytest.mark.parametrize('limit', ['32', '0'])
def test_contraction_direct_write_path(limit, tmp_path):
    """Chunks with more than 64 local DOFs per side take a direct-write path in the contraction that ordinary
    meshes never reach; the TEST build of the library (libthincurr_b200_test.so, -DTW_TEST_HOOKS) reads
    THINCURR_B200_DRAIN_LIMIT to lower the limit so that the path runs (rows, columns, both) and must reproduce the
    normal build bit for bit, for the single-device and the row-sharded (mirror-writing) tiles.  Runs in a
    subprocess because the product library has no such switch."""
    import subprocess
    import sys
    from openfusiontoolkit_b200.build import build_test
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'drain.py'
    script.write_text(_DRAIN_SCRIPT)
    env = dict(os.environ, THINCURR_B200_LIB=build_test())
    res = subprocess.run([sys.executable, str(script), root, limit], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and 'DRAIN_OK' in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
