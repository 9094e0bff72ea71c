"""Drop-in boundary on the GPU: the reference-shaped call paths end to end.
 * goldens through `setup_model(mesh_file=<reference fixture .h5>, xml_filename=...)` -> compute_Lmat / compute_Rmat ->
   get_eigs (Lanczos on the device-resident matrix) -- the flow of the reference's own test (test_ThinCurr.py:182-203);
 * the UNMODIFIED reference Python package driving libthincurr_b200.so through that flow (when staged under
   baseline/_ref by tools/install_reference_python.py);
 * the Fortran-host pair thincurr_b200_model_from_tw + thincurr_b200_Lmat_host with Fortran-convention arrays;
 * thincurr_apply_Lmat / thincurr_eigenvalues against numpy / scipy;
 * byte-level checks of the operator caches against the record layout thin_wall.F90 writes
   (:1161-1183 Lmat.save and mutual, :755-763 Mcoil.save, :1675-1684 Msen.save), read here by an independent
   pure-Python reader of gfortran's unformatted-sequential framing."""
import ctypes
import os
import re
import struct
import numpy as np
import pytest
from helpers import GOLDEN, MU0, goldens, load_mesh, split_nodesets, ref_circle, ref_floop
from oracle import tw_oracle as tw
import _ref_layer

pytestmark = pytest.mark.gpu
G = goldens()


@pytest.fixture(scope='module')
def env():
    from openfusiontoolkit_b200 import OFT_env
    return OFT_env(nthreads=-1)


def _xml(tmp_path, eta):
    p = tmp_path / 'oft_in.xml'
    p.write_text('<oft>\n  <thincurr>\n    <eta>%.6E</eta>\n  </thincurr>\n</oft>\n' % eta)
    return str(p)


def fortran_records(path):
    """Records of a gfortran unformatted-sequential file (4-byte length markers before and after each record;
    negative markers chain sub-records of records beyond 2 GiB)."""
    out = []
    with open(path, 'rb') as f:
        data = f.read()
    pos, cur = 0, b''
    while pos < len(data):
        head = struct.unpack_from('<i', data, pos)[0]
        n = abs(head)
        body = data[pos + 4:pos + 4 + n]
        tail = struct.unpack_from('<i', data, pos + 4 + n)[0]
        assert abs(tail) == n
        cur += body
        pos += 8 + n
        if head >= 0:
            out.append(cur)
            cur = b''
    assert cur == b''
    return out


@pytest.mark.parametrize('name,js', [('plate', 0), ('cyl', 2), ('torus', 0)])
def test_eigen_goldens_through_the_mesh_file_path(env, name, js, tmp_path):
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    g = G['eig_' + name]
    T = ThinCurr(env)
    T.setup_model(mesh_file=os.path.join(GOLDEN, 'ref_h5', 'tw_test-%s.h5' % name), xml_filename=_xml(tmp_path, 10.0 * MU0),
                  jumper_start=js)
    cache = str(tmp_path / 'Lmat.save')
    T.compute_Lmat(cache_file=cache)
    T.compute_Rmat()
    vals, vecs = T.get_eigs(4)
    assert np.abs(vals / np.array(g['vals']) - 1.0).max() < g['tol']
    # Lanczos on the device against the dense generalised eigen solve: 1e-8 (north star)
    import scipy.linalg as sl
    w = np.sort(sl.eigh(np.array(T.Lmat), T.Rmat.toarray(), eigvals_only=True))[::-1][:4]
    assert np.abs(vals / w - 1.0).max() < 1e-8
    for k in range(4):
        r = T.Lmat @ vecs[k] - vals[k] * (T.Rmat @ vecs[k])
        assert np.linalg.norm(r) < 1e-7 * np.linalg.norm(T.Lmat @ vecs[k])
    # the cache written by that call has the reference's layout and is read back by a second model
    recs = fortran_records(cache)
    N = T.nelems
    assert len(recs) == N + 1
    hl, hr = T.model_hashes()
    assert struct.unpack('<6i', recs[0]) == (N, T.nc, hl, hl, hr, hr)
    for i in (0, 1, N // 2, N - 1):
        assert np.array_equal(np.frombuffer(recs[i + 1]), T.Lmat[i, i:])
    T2 = ThinCurr(env)
    T2.setup_model(mesh_file=os.path.join(GOLDEN, 'ref_h5', 'tw_test-%s.h5' % name), xml_filename=_xml(tmp_path, 10.0 * MU0),
                   jumper_start=js)
    T2.compute_Lmat(cache_file=cache)
    assert np.array_equal(T2.Lmat, T.Lmat)


def test_hash_is_over_the_one_based_oriented_connectivity(env):
    """oft_simple_hash(C_LOC(mesh%lc), 4*3*nc) (thin_wall.F90:920): the Fortran array holds 1-based, orientation-synced
    vertex ids; the known-answer hash is recomputed here in pure Python (oft_local_c.c:86-98, Jenkins one-at-a-time)."""
    from openfusiontoolkit_b200.ThinCurr import ThinCurr

    def oat(b):
        h = 0
        for x in b:
            h = (h + x) & 0xffffffff
            h = (h + (h << 10)) & 0xffffffff
            h ^= h >> 6
        h = (h + (h << 3)) & 0xffffffff
        h ^= h >> 11
        h = (h + (h << 15)) & 0xffffffff
        return h - (1 << 32) if h & 0x80000000 else h
    m = load_mesh('cyl')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=split_nodesets(m, 2))
    A = T.get_model_arrays()
    hl, hr = T.model_hashes()
    assert hl == oat((A['lc'] + 1).astype('<i4').tobytes())
    assert hr == oat(np.ascontiguousarray(m['r'], '<f8').tobytes())


def test_fortran_host_pair_reproduces_thincurr_Lmat(env):
    """thincurr_b200_model_from_tw (arrays of a Fortran tw_type) + thincurr_b200_Lmat_host
    (caller-owned Lmat(nelems,nelems)) == thincurr_Lmat, bit for bit."""
    from openfusiontoolkit_b200 import _interface as I
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('torus')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0])
    T.compute_Lmat()
    A = T.get_model_arrays()
    lfh1 = A['lfh'].copy()
    lfh1[:, 1] += 1
    tw_ptr = ctypes.c_void_p()
    rc = I.b200_model_from_tw(T.np, np.ascontiguousarray(m['r'], np.float64), T.nc, np.ascontiguousarray(A['lc'] + 1, np.int32), None,
                              np.ascontiguousarray(A['pmap'], np.int32), T.np_active, T.nholes,
                              np.ascontiguousarray(A['kfh'] + 1, np.int32), np.ascontiguousarray(lfh1, np.int32).ctypes.data_as(ctypes.c_void_p),
                              np.ascontiguousarray(A['ca']).ctypes.data_as(ctypes.c_void_p),
                              np.ascontiguousarray(A['qbasis']).ctypes.data_as(ctypes.c_void_p), ctypes.byref(tw_ptr))
    assert rc == 0, I.b200_last_error()
    L = np.full((T.nelems, T.nelems), np.nan)   # pageable caller memory, like a Fortran ALLOCATE
    assert I.b200_Lmat_host(tw_ptr, L) == 0, I.b200_last_error()
    assert np.array_equal(L, T.Lmat)
    I.b200_destroy(tw_ptr)


def test_apply_Lmat(env):
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('ex_torus')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0])
    T.compute_Lmat()
    x = np.random.default_rng(5).normal(size=T.nelems)
    y = T.apply_Lmat(x)
    ref = T.Lmat @ x
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(T.Lmat).sum(axis=1).max() * np.abs(x).max()


def test_entry_points_outside_the_backend_report_through_error_str(env):
    from openfusiontoolkit_b200 import _interface as I
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('plate')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'])
    err = env.get_c_errorbuff()
    I.thincurr_freq_response(T.tw_obj, False, 0, 1.e3, np.zeros((2, T.nelems)), None, err)
    assert b'not provided' in err.value
    err = env.get_c_errorbuff()
    I.thincurr_eigenvalues(T.tw_obj, False, 2, np.zeros(2), np.zeros((2, T.nelems)), None, err)
    assert err.value == b'Inductance matrix required, but not computed'   # thincurr_f.F90:989-992
    T.compute_Lmat()
    err = env.get_c_errorbuff()
    I.thincurr_eigenvalues(T.tw_obj, False, 2, np.zeros(2), np.zeros((2, T.nelems)), None, err)
    assert err.value == b'Resistance matrix required, but not computed'   # :993-996
    th = np.zeros(T.nregs)
    err = env.get_c_errorbuff()
    I.thincurr_get_thickness(T.tw_obj, th, err)
    assert err.value == b''
    v = np.ones(T.np)
    I.thincurr_scale_va(T.tw_obj, v, False)
    assert np.isclose(v.sum(), T.get_model_arrays()['ca'].sum(), rtol=1e-12)   # vertex areas sum to the surface area


def test_coil_sensor_and_mutual_cache_files(env, tmp_path):
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    from openfusiontoolkit_b200.ThinCurr.sensor import circular_flux_loop, save_sensors
    g = G['fr_passive']
    m = load_mesh('plate')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'])
    T.set_coils('vcoil', [[dict(pts=ref_circle(R, Z), radius=1e-2, res_per_len=1.256637e-5)] for (R, Z) in g['vcoils']])
    T.set_coils('icoil', [[dict(pts=ref_circle(R, Z))] for (R, Z) in g['icoils']])
    N, nv, ni = T.nelems, T.n_vcoils, T.n_icoils
    assert (nv, ni) == (1, 1)
    mc_file = str(tmp_path / 'Mcoil.save')
    Mc = np.array(T.compute_Mcoil(cache_file=mc_file))
    recs = fortran_records(mc_file)
    assert len(recs) == 3 and struct.unpack('<3i', recs[0]) == (N, nv, ni)
    assert len(recs[1]) == 8 * N * nv and np.array_equal(np.frombuffer(recs[2]), Mc.ravel())   # Ael2dr(nelems,n_icoils)
    sens = [circular_flux_loop(R, Z, 'FLOOP_%d' % k) for k, (R, Z) in enumerate(g['floops'])]
    floops = str(tmp_path / 'floops.loc')
    save_sensors(sens, floops)
    ms_file = str(tmp_path / 'Msen.save')
    Ms, Msc, _ = T.compute_Msensor(sensor_file=floops, cache_file=ms_file)
    Ms, Msc = np.array(Ms), np.array(Msc)
    recs = fortran_records(ms_file)
    assert len(recs) == 3 and struct.unpack('<4i', recs[0]) == (N, nv, ni, 2)
    assert np.array_equal(np.frombuffer(recs[1]), Ms.ravel()) and np.array_equal(np.frombuffer(recs[2]), Msc.ravel())
    # both caches are read back instead of rebuilt (second model, same sizes)
    T2 = ThinCurr(env)
    T2.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'])
    T2.set_coils('vcoil', [[dict(pts=ref_circle(R, Z), radius=1e-2, res_per_len=1.256637e-5)] for (R, Z) in g['vcoils']])
    T2.set_coils('icoil', [[dict(pts=ref_circle(R, Z))] for (R, Z) in g['icoils']])
    assert np.array_equal(T2.compute_Mcoil(cache_file=mc_file), Mc)
    Ms2, Msc2, _ = T2.compute_Msensor(sensor_file=floops, cache_file=ms_file)
    assert np.array_equal(Ms2, Ms) and np.array_equal(Msc2, Msc)
    # mutual cache: header (col nelems, row nelems, hashes), one record per row-model element
    m2 = load_mesh('passive')
    T3 = ThinCurr(env)
    T3.setup_model(r=m2['r'] + np.array([0.0, 0.0, 0.3]), lc=m2['lc'], reg=m2['reg'])
    mu_file = str(tmp_path / 'Mutual.save')
    M = T.cross_coupling(T3, cache_file=mu_file)
    recs = fortran_records(mu_file)
    h1, h3 = T.model_hashes(), T3.model_hashes()
    assert struct.unpack('<6i', recs[0]) == (T3.nelems, T.nelems, h1[0], h3[0], h1[1], h3[1])
    assert len(recs) == T.nelems + 1 and np.array_equal(np.frombuffer(recs[5]), M[4])
    M2 = T.cross_coupling(T3, cache_file=mu_file)
    assert np.array_equal(M2, M)


def test_streamed_lmat_save_matches_the_in_memory_writer(env, tmp_path):
    """thincurr_b200_Lmat_save_begin/_rows (shards written at their file offsets) produce the same bytes as the
    in-memory cache writer of thincurr_Lmat."""
    import torch
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('ex_torus')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0])
    a = str(tmp_path / 'a.save')
    T.compute_Lmat(cache_file=a)
    b = str(tmp_path / 'b.save')
    T.save_Lmat_begin(b)
    for s in (2, 0, 1):
        rows = T.shard_rows(3, s)
        out = torch.empty((len(rows), T.nelems), dtype=torch.float64, device='cuda')
        T.compute_Lmat_shard(3, s, out)
        torch.cuda.synchronize()
        T.save_Lmat_rows(b, 3, s, False, out.data_ptr(), T.nelems)
    assert open(a, 'rb').read() == open(b, 'rb').read()
    # and the streamed host export lands the rows in the reference layout
    full = np.zeros((T.nelems, T.nelems))
    for s in range(3):
        rows = T.shard_rows(3, s)
        out = torch.empty((len(rows), T.nelems), dtype=torch.float64, device='cuda')
        T.compute_Lmat_shard(3, s, out)
        torch.cuda.synchronize()
        T.rows_to_host(3, s, False, out.data_ptr(), T.nelems, full)
    assert np.array_equal(full, T.Lmat)


@pytest.mark.skipif(not _ref_layer.available(), reason='reference Python layer not staged (tools/install_reference_python.py)')
@pytest.mark.parametrize('name,js', [('plate', 0), ('cyl', 2), ('torus', 0)])
def test_unmodified_reference_python_layer_reaches_the_goldens(name, js, tmp_path):
    """The reference's OWN ThinCurr class (sources untouched) on libthincurr_b200.so: the eigenvalue test of
    src/tests/physics/test_ThinCurr.py:182-203,983-986,1028-1031,1075-1078 minus setup_io (plot files)."""
    g = G['eig_' + name]
    body = r'''
env = OFT_env(nthreads=2, quiet=True)
tw = ThinCurr(env)
tw.setup_model(mesh_file=os.path.join(GOLDEN, 'ref_h5', 'tw_test-%s.h5'), xml_filename=%r, jumper_start=%d)
tw.compute_Mcoil()
tw.compute_Lmat(cache_file=%r)
tw.compute_Rmat()
vals, vecs = tw.get_eigs(4, direct=False)
print('EIGS', ' '.join('%%.9e' %% v for v in vals))
x = np.linspace(0.0, 1.0, tw.nelems)
print('APPLY', float(np.abs(tw.Lmat @ x - tw.apply_Lmat(x)).max() / np.abs(tw.Lmat @ x).max()))
''' % (name, _xml(tmp_path, 10.0 * MU0), js, str(tmp_path / 'Lmat.save'))
    res = _ref_layer.run(body)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    vals = np.array([float(v) for v in re.search(r'EIGS (.*)', res.stdout).group(1).split()])
    assert np.abs(vals / np.array(g['vals']) - 1.0).max() < g['tol']
    assert float(re.search(r'APPLY (\S+)', res.stdout).group(1)) < 1e-13
