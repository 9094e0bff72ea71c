"""Drop-in boundary, CPU side (no GPU compute is called):
 * the library exports every name the reference's Python layer binds at import
   (src/python/OpenFUSIONToolkit/ThinCurr/_interface.py:17-119 and OpenFUSIONToolkit/_interface.py:98-132);
 * the UNMODIFIED reference Python package imports against it and drives model setup from the reference's own native
   mesh files + an oft_in.xml (the normal entry of the reference, thincurr_f.F90:49-228), reproducing the model the
   in-memory path and the oracle build;
 * a Fortran host's arrays (1-based kfh, lfh(2,:)) rebuild the same model through thincurr_b200_model_from_tw."""
import ctypes
import json
import os
import re
import numpy as np
import pytest
from helpers import GOLDEN, MU0, load_mesh, split_nodesets
from oracle import tw_oracle as tw
import _ref_layer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_IF = '/root/reference/src/python/OpenFUSIONToolkit'

XML = '''<oft>
  <thincurr>
    <eta>%s</eta>
    <icoils>
      <coil_set><coil scale="1.0">1.5, 0.5</coil></coil_set>
      <coil_set><coil>1.5, -0.5</coil></coil_set>
    </icoils>
    <vcoils>
      <coil_set res_per_len="1.E-4" radius="1.E-2"><coil>0.6, 0.0</coil></coil_set>
    </vcoils>
  </thincurr>
</oft>
'''


@pytest.fixture(scope='module')
def env():
    from openfusiontoolkit_b200 import OFT_env
    return OFT_env(nthreads=-1)


def test_library_exports_every_name_the_reference_binds():
    names = json.load(open(os.path.join(GOLDEN, 'ref_interface_symbols.json')))['names']
    base = ['oftpy_init', 'oftpy_set_debug', 'oftpy_set_nthreads', 'oftpy_load_xml', 'oft_setup_smesh', 'oft_smesh_get',
            'oft_setup_vmesh', 'oft_vmesh_get', 'dump_cov']
    if os.path.isdir(REF_IF):  # the fixture list is in step with the reference's sources
        src = open(os.path.join(REF_IF, 'ThinCurr', '_interface.py')).read()
        assert sorted(set(re.findall(r'oftpy_lib\.(\w+)', src))) == names
        src = open(os.path.join(REF_IF, '_interface.py')).read()
        assert sorted(set(re.findall(r'oftpy_lib\.(\w+)', src))) == sorted(base)
    assert len(names) == 26
    lib = ctypes.CDLL(os.path.join(ROOT, 'openfusiontoolkit_b200', 'libthincurr_b200.so'))
    missing = [n for n in names + base if not hasattr(lib, n)]
    assert not missing, 'bound by the reference at import but not exported: %s' % missing


def test_product_library_has_no_test_hooks():
    lib = ctypes.CDLL(os.path.join(ROOT, 'openfusiontoolkit_b200', 'libthincurr_b200.so'))
    for n in ('thincurr_b200_probe_pairs', 'thincurr_b200_probe_phipot', 'thincurr_b200_probe_rsqrt'):
        assert not hasattr(lib, n)
    blob = open(os.path.join(ROOT, 'openfusiontoolkit_b200', 'libthincurr_b200.so'), 'rb').read()
    assert b'THINCURR_B200_DEBUG_SKIP' not in blob and b'THINCURR_B200_DRAIN_LIMIT' not in blob


@pytest.mark.parametrize('name,js', [('plate', 0), ('cyl', 2), ('torus', 0)])
def test_setup_from_reference_mesh_file_and_xml(env, name, js, tmp_path):
    """setup_model(mesh_file=..., xml_filename=...) on the reference's own fixture files == the in-memory path == oracle."""
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    xml = tmp_path / 'oft_in.xml'
    xml.write_text(XML % '1.257E-5')
    T = ThinCurr(env)
    T.setup_model(mesh_file=os.path.join(GOLDEN, 'ref_h5', 'tw_test-%s.h5' % name), xml_filename=str(xml), jumper_start=js)
    m = load_mesh(name)
    ns = split_nodesets(m, js)
    cl = m['sidesets'][0] if m['sidesets'] else None
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=ns, closures=cl if cl is not None else (), eta=[1.257e-5])
    assert (T.np, T.nc, T.np_active, T.nholes) == (O.np_, O.nc, O.np_active, O.nholes)
    assert T.n_icoils == 2 and T.n_vcoils == 1 and T.nelems == O.np_active + O.nholes + 1
    A = T.get_model_arrays()
    assert np.array_equal(A['pmap'], O.pmap) and np.array_equal(A['lc'], O.lc)
    assert np.array_equal(A['kfh'], O.kfh) and np.array_equal(A['lfh'], O.lfh.reshape(-1, 2))
    assert np.allclose(T.get_eta_values(), 1.257e-5, rtol=1e-14)
    T2 = ThinCurr(env)
    T2.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=ns if len(ns) else None, closures=cl)
    assert T.model_hashes() == T2.model_hashes()


def test_model_from_fortran_host_arrays(env):
    """thincurr_b200_model_from_tw: the arrays of a Fortran tw_type (oriented 1-based lc, pmap, 1-based kfh,
    lfh(2,:) with 1-based local vertex) give the model `setup_model` builds (no setup work repeated)."""
    from openfusiontoolkit_b200 import _interface as I
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m = load_mesh('torus')
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0])
    A = T.get_model_arrays()
    lfh1 = A['lfh'].copy()
    lfh1[:, 1] += 1
    tw_ptr = ctypes.c_void_p()
    rc = I.b200_model_from_tw(T.np, np.ascontiguousarray(m['r'], np.float64), T.nc, np.ascontiguousarray(A['lc'] + 1, np.int32), None,
                              np.ascontiguousarray(A['pmap'], np.int32), T.np_active, T.nholes,
                              np.ascontiguousarray(A['kfh'] + 1, np.int32), np.ascontiguousarray(lfh1, np.int32).ctypes.data_as(ctypes.c_void_p),
                              None, None, ctypes.byref(tw_ptr))
    assert rc == 0, I.b200_last_error()
    pm, lc, kfh = np.zeros(T.np, np.int32), np.zeros((T.nc, 3), np.int32), np.zeros(T.nc + 1, np.int32)
    qb, ca = np.zeros((T.nc, 3, 3)), np.zeros(T.nc)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    nfh = I.b200_get_model(tw_ptr, vp(pm), vp(lc), vp(kfh), None, vp(qb), vp(ca))
    lfh = np.zeros((max(nfh, 1), 2), np.int32)
    I.b200_get_model(tw_ptr, None, None, None, vp(lfh), None, None)
    assert np.array_equal(pm, A['pmap']) and np.array_equal(lc, A['lc']) and np.array_equal(kfh, A['kfh'])
    assert np.array_equal(lfh[:nfh], A['lfh'])
    assert np.array_equal(qb, A['qbasis']) and np.array_equal(ca, A['ca'])
    I.b200_destroy(tw_ptr)


@pytest.mark.skipif(not _ref_layer.available(), reason='reference Python layer not staged (tools/install_reference_python.py)')
def test_unmodified_reference_python_layer_loads_and_sets_up(tmp_path):
    """`import OpenFUSIONToolkit.ThinCurr` (reference sources, untouched) binds every symbol from libthincurr_b200.so; its
    ThinCurr.setup_model(mesh_file, xml) / compute_Rmat / get_eta_values run on it; entry points outside this backend raise
    through the reference's own error convention."""
    xml = tmp_path / 'oft_in.xml'
    xml.write_text(XML % '1.257E-5')
    body = r'''
import torch
env = OFT_env(nthreads=2, quiet=True)
tw = ThinCurr(env)
tw.setup_model(mesh_file=os.path.join(GOLDEN, 'ref_h5', 'tw_test-torus.h5'), xml_filename=%r)
print('SIZES', tw.np, tw.nc, tw.np_active, tw.nholes, tw.n_vcoils, tw.n_icoils, tw.nelems)
tw.compute_Rmat()
print('RMAT', tw.Rmat.shape, tw.Rmat.nnz, float(abs(tw.Rmat - tw.Rmat.T).max()))
print('ETA', tw.get_eta_values()[0])
try:
    tw.compute_freq_response(fdriver=np.zeros((2, tw.nelems)), freq=1.e3)
    print('FR no error')
except Exception as e:
    print('FR_ERR', e)
if not torch.cuda.is_available():
    try:
        tw.compute_Lmat()
        print('LMAT no error')
    except Exception as e:
        print('LMAT_ERR', e)
    try:
        tw.compute_Mcoil()
        print('MCOIL no error')
    except Exception as e:
        print('MCOIL_ERR', e)
''' % str(xml)
    res = _ref_layer.run(body)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    out = res.stdout
    m = load_mesh('torus')
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=m['nodesets'], closures=m['sidesets'][0], eta=[1.257e-5])
    sizes = [int(v) for v in re.search(r'SIZES (.*)', out).group(1).split()]
    assert sizes == [O.np_, O.nc, O.np_active, O.nholes, 1, 2, O.np_active + O.nholes + 1]
    assert 'FR_ERR thincurr_freq_response is not provided by the B200 operator-build backend' in out
    assert 'LMAT no error' not in out and 'MCOIL no error' not in out
    if 'LMAT_ERR' in out:  # no GPU here: the reference's own precondition (thincurr_f.F90:555-558), then "no CPU fallback"
        assert 'LMAT_ERR Coil mutuals required' in out and re.search(r'MCOIL_ERR .*CUDA', out)
    assert abs(float(re.search(r'ETA (\S+)', out).group(1)) / 1.257e-5 - 1) < 1e-12
