"""CPU-only checks of the host side of the B200 library (no GPU compute is called):
the shared library loads and exports every symbol include/thincurr_b200.h declares, the C++ model
setup (tw_setup restatement) agrees exactly with the oracle's, the resistance matrix and hashes
agree, and the owner-computes plan partitions rows and tiles correctly."""
import os
import re
import ctypes
import numpy as np
import pytest
from helpers import MU0, load_mesh, split_nodesets, dummy_mesh
from oracle import tw_oracle as tw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MESHES = [('plate', 0), ('cyl', 2), ('torus', 0), ('ex_cyl', -1), ('ex_torus', 0), ('ex_ports', 0)]


@pytest.fixture(scope='module')
def env():
    from openfusiontoolkit_b200 import OFT_env
    return OFT_env(nthreads=-1)


def _pair(env, name, js, eta=10.0):
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh(name)
    ns = split_nodesets(m, js)
    cl = m['sidesets'][0] if m['sidesets'] else None
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=ns, closures=cl if cl is not None else (), eta=[eta * MU0])
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=ns if len(ns) else None, closures=cl)
    return O, T


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'thincurr_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = set(re.findall(r'\b((?:thincurr|oftpy)_\w+)\s*\(', hdr))
    assert len(names) >= 30
    lib = ctypes.CDLL(os.path.join(ROOT, 'openfusiontoolkit_b200', 'libthincurr_b200.so'))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, 'declared but not exported: %s' % missing


def test_fortran_interface_matches_header():
    """Every procedure bound in the ISO_C_BINDING interface module exists in the C header."""
    f90 = open(os.path.join(ROOT, 'include', 'thincurr_b200_f.F90')).read()
    hdr = open(os.path.join(ROOT, 'include', 'thincurr_b200.h')).read()
    bound = re.findall(r'BIND\(C,\s*NAME="(\w+)"\)', f90)
    assert len(bound) >= 8
    for n in bound:
        assert re.search(r'\b%s\s*\(' % n, hdr), n


@pytest.mark.parametrize('name,js', MESHES)
def test_setup_matches_oracle(env, name, js):
    O, T = _pair(env, name, js)
    assert (T.np, T.nc, T.np_active, T.nholes, T.nelems) == (O.np_, O.nc, O.np_active, O.nholes, O.nelems)
    A = T.get_model_arrays()
    assert np.array_equal(A['pmap'], O.pmap)
    assert np.array_equal(A['lc'], O.lc), 'orientation sync must flip the same cells'
    assert np.array_equal(A['kfh'], O.kfh)
    assert np.array_equal(A['lfh'], np.asarray(O.lfh).reshape(-1, 2)[:O.nfh])
    assert np.abs(A['qbasis'] - O.qbasis).max() <= 1e-13 * np.abs(O.qbasis).max()
    assert np.abs(A['ca'] - O.ca).max() <= 1e-15 * O.ca.max()
    assert T.model_hashes() == (O.hash_lc(), O.hash_r())


@pytest.mark.parametrize('name,js', [('plate', 0), ('cyl', 2), ('torus', 0)])
def test_rmat_matches_oracle(env, name, js):
    O, T = _pair(env, name, js)
    T.set_eta_values(eta_surf=np.array([10.0 * MU0]))
    T.compute_Rmat()
    Ro = O.compute_Rmat().toarray()
    Rg = T.Rmat.toarray()
    assert np.abs(Rg - Ro).max() <= 1e-13 * np.abs(Ro).max()
    assert np.allclose(T.get_eta_values(), 10.0 * MU0, rtol=1e-15)


def test_error_conventions(env):
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    T = ThinCurr(env)
    with pytest.raises(ValueError):
        T.setup_model()
    with pytest.raises(Exception):
        T.setup_model(mesh_file='/nonexistent/mesh.h5')
    m = load_mesh('plate')
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'])
    with pytest.raises(ValueError):
        T.setup_model(r=m['r'], lc=m['lc'])
    with pytest.raises(IndexError):
        T.set_eta_values(eta_surf=np.ones(3))
    with pytest.raises(ValueError):
        T.set_eta_values(eta_surf=-np.ones(1))
    with pytest.raises(Exception, match='HODLR'):
        T.compute_Lmat(use_hodlr=True)
    bad = m['lc'].copy()
    bad[3, 1] = m['r'].shape[0] + 5      # malformed connectivity is refused, not read out of bounds
    with pytest.raises(Exception, match='outside'):
        ThinCurr(env).setup_model(r=m['r'], lc=bad, reg=m['reg'])
    with pytest.raises(Exception, match='outside'):
        ThinCurr(env).setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=[np.array([0, 1, 10 ** 6])])


def test_no_cpu_fallback(env):
    """Without a CUDA device every operator build must fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('plate')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'])
    with pytest.raises(Exception, match='CUDA'):
        T.compute_Lmat()
    with pytest.raises(Exception, match='CUDA'):
        T.compute_Bmat()


@pytest.mark.parametrize('name,js', [('torus', 0), ('ex_torus', 0), ('ex_ports', 0)])
@pytest.mark.parametrize('nshards', [1, 2, 8])
def test_plan_partitions_rows(env, name, js, nshards):
    O, T = None, None
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh(name)
    ns = split_nodesets(m, js)
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=ns if len(ns) else None,
                  closures=m['sidesets'][0] if m['sidesets'] else None)
    rows = [T.shard_rows(nshards, s) for s in range(nshards)]
    allr = np.concatenate(rows)
    assert np.array_equal(np.sort(allr), np.arange(T.nelems)), 'shards must partition the rows'
    info = T.plan_info()
    assert info['patch_cells'] >= T.nc and info['ntiles'] == info['npatch'] * (info['npatch'] + 1) // 2
    if nshards > 1 and T.nelems > 2000:
        sizes = np.array([len(r) for r in rows], float)
        assert sizes.max() / sizes.mean() < 1.35, 'row blocks are balanced by cell count'


def test_dummy_mesh_generator_matches_reference_counts():
    from openfusiontoolkit_b200.ThinCurr.meshing import build_ThinCurr_dummy, build_torus_vessel
    r, lc = build_ThinCurr_dummy([0., 0., 0.], size=0.25, nsplit=1)
    r2, lc2 = dummy_mesh([0., 0., 0.], size=0.25, nsplit=1)
    assert np.array_equal(lc, lc2) and np.allclose(r, r2)
    m = build_torus_vessel(24, 48, nports=4)
    used = np.zeros(len(m['r']), bool)
    used[m['lc'].ravel()] = True
    assert used.all() and len(m['nodesets']) == 2 + 4


def test_h5_writer_roundtrip(tmp_path):
    """The library's minimal HDF5 writer (container of the Bmat cache, thin_wall.F90:2208-2225): the file parses
    with the oracle-side reader that was validated on the reference's libhdf5-written fixtures, datasets come back
    bit for bit (float64 2-D, int32 1-D, zero-size), and the metadata layout is the one libhdf5 produces."""
    import ctypes
    import struct
    from openfusiontoolkit_b200 import _interface as I
    from oracle import h5min
    rng = np.random.default_rng(3)
    A = rng.standard_normal((37, 53))
    H = np.array([659, 1440, -123456789, 2 ** 31 - 1], dtype=np.int32)
    Z = np.zeros((0, 37))
    names = [b'Bel_X', b'MODEL_hash', b'Bdr_X', b'Bel_Y']
    arrs = [A, H, Z, np.ascontiguousarray(A.T)]
    n = len(names)
    dims = [37, 53, 4, 0, 37, 53, 37]
    fn = str(tmp_path / 'bmat.h5')
    rc = I.b200_h5_write(fn.encode(), n, (ctypes.c_char_p * n)(*names), (ctypes.c_int * n)(1, 0, 1, 1), (ctypes.c_int * n)(2, 1, 2, 2),
                         (ctypes.c_int64 * len(dims))(*dims), (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs]))
    assert rc == 0
    h = h5min.H5(fn)
    t = h.tree()
    assert sorted(t) == ['Bdr_X', 'Bel_X', 'Bel_Y', 'MODEL_hash']
    assert np.array_equal(h.read(t['Bel_X']), A) and np.array_equal(h.read(t['Bel_Y']), A.T)
    assert np.array_equal(h.read(t['MODEL_hash']), H) and h.read(t['MODEL_hash']).dtype == np.int32
    assert h.read(t['Bdr_X']).shape == (0, 37)
    raw = open(fn, 'rb').read()
    # superblock v0: 8-byte offsets/lengths, group K 4/16, end-of-file address = file size, root entry cached as a group
    assert raw[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0]) and struct.unpack_from('<HH', raw, 16) == (4, 16)
    assert struct.unpack_from('<Q', raw, 40)[0] == len(raw) and struct.unpack_from('<I', raw, 72)[0] == 1
    bt, heap = struct.unpack_from('<QQ', raw, 80)
    assert raw[bt:bt + 4] == b'TREE' and raw[heap:heap + 4] == b'HEAP' and heap - bt == 544


def test_stream_plan_bands(env, monkeypatch):
    """Banded plan of the streamed single-device build (thincurr_b200_stream_plan, host-only): equal ranges of reference
    ids, every band's patches hold exactly the DOFs of its range (the hole DOFs in the last band), small models and a
    reference numbering without locality are built the ordinary way."""
    import bench
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    from openfusiontoolkit_b200.ThinCurr.meshing import build_torus_vessel
    from openfusiontoolkit_b200 import _interface as I
    m = load_mesh('ex_torus')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0])
    assert T.stream_plan() is None   # 2 395 DOFs: one launch, one copy
    m = bench.make_mesh('vessel20k')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], nodesets=m['nodesets'], closures=m['closures'])
    ref, pat = T.stream_plan()
    nb = len(ref) - 1
    assert 2 <= nb <= 10 and ref[0] == 0 and ref[-1] == T.nelems and pat[0] == 0
    assert (np.diff(ref) > 0).all() and (np.diff(pat) > 0).all()
    widths = np.diff(ref[:-1])
    assert widths.max() - widths.min() <= 64   # equal ranges (cut at multiples of 32)
    patch_of = np.zeros(T.nelems, dtype=np.int32)
    assert I.b200_dof_patches(T.tw_obj, 1, patch_of) == 0
    for b in range(nb):
        p = patch_of[ref[b]:ref[b + 1]]
        assert p.min() >= pat[b] and p.max() < pat[b + 1], 'band %d: patches outside its range' % b
    assert pat[-1] == T.plan_info()['npatch']
    assert (patch_of[T.np_active:] >= pat[-2]).all()   # hole patches belong to the last band
    # the same vessel with a random vertex numbering: ranges of reference ids are scattered over the surface
    nt, nphi = bench.WORKLOADS['vessel20k']
    mp = build_torus_vessel(nt, nphi, nports=10, permute_seed=3)
    T = ThinCurr(env)
    T.setup_model(r=mp['r'], lc=mp['lc'], nodesets=mp['nodesets'], closures=mp['closures'])
    assert T.stream_plan() is None
