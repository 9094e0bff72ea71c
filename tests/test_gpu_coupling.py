"""GPU parity of the coil / sensor coupling and B-field reconstruction operators and of the mutual
(cross-coupling) inductance against the oracle (tw_compute_Ael2dr/_Lmat_coils :567-883,
tw_compute_mutuals :1418-1686, tw_compute_Bops :1989-2226, tw_compute_LmatDirect with col_model),
plus the reference's passive-V-coil eigenvalue golden and the frequency-response goldens computed
from the GPU-built operators."""
import numpy as np
import pytest
from helpers import MU0, dummy_mesh, goldens, load_mesh, mutual_abs_sum, split_nodesets, ref_circle, ref_floop
from oracle import tw_oracle as tw

pytestmark = pytest.mark.gpu
G = goldens()


def relerr(A, B):
    return np.abs(np.asarray(A) - np.asarray(B)).max() / np.abs(B).max()


@pytest.fixture(scope='module')
def env():
    from openfusiontoolkit_b200 import OFT_env
    return OFT_env(nthreads=-1)


def make(env, name, g, vcoils=None, icoils=None):
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    if name == 'passive':
        r, lc = dummy_mesh([0.0, 0.0, 10.0], size=0.25, nsplit=1)
        m = dict(r=r, lc=lc, reg=None, nodesets=[], sidesets=[])
    else:
        m = load_mesh(name)
    ns = split_nodesets(m, g.get('jumper_start', 0))
    cl = m['sidesets'][0] if m['sidesets'] else None
    eta = g.get('eta', 10.0) * MU0
    vc = [[dict(pts=ref_circle(R, Z), scale=1.0, radius=1.e-2, res_per_len=1.256637E-5)] for (R, Z) in (vcoils or [])]
    ic = [[dict(pts=ref_circle(R, Z), scale=1.0) for (R, Z) in icoils]] if icoils else []
    O = tw.OracleModel(m['r'], m['lc'], m['reg'], nodesets=ns, closures=cl if cl is not None else (), eta=[eta],
                       vcoils=tw.CoilSets([dict(filaments=[(f['pts'], 1.0, 1.e-2, 1.256637E-5) for f in s]) for s in vc]),
                       icoils=tw.CoilSets([dict(filaments=[(f['pts'], 1.0, -1.0, -1.0) for f in s]) for s in ic]))
    T = ThinCurr(env)
    T.setup_model(r=m['r'], lc=m['lc'], reg=m['reg'], nodesets=ns if len(ns) else None, closures=cl)
    if vc:
        T.set_coils('vcoil', vc)
    if ic:
        T.set_coils('icoil', ic)
    T.set_eta_values(eta_surf=np.array([eta]))
    return O, T


def test_passive_vcoil_eigen_golden(env):
    """run_eig 'passive' case (test_ThinCurr.py:1163-1168): V-coil rows of L and R."""
    import scipy.linalg as sl
    g = G['eig_passive']
    O, T = make(env, 'passive', g, vcoils=g['vcoils'])
    Mo = O.compute_Mcoil()
    Mg = T.compute_Mcoil()
    assert Mg.shape == Mo.shape
    Lo = O.compute_Lmat()
    T.compute_Lmat()
    assert relerr(T.Lmat, Lo) < 1e-12
    T.compute_Rmat()
    w = np.sort(np.abs(sl.eigh(T.Lmat, T.Rmat.toarray(), eigvals_only=True)))[::-1][:4]
    assert np.abs(w / np.array(g['vals']) - 1.0).max() < g['tol']


def test_vcoil_precondition(env):
    """thincurr_f.F90:555-558: L of a model with V-coils needs the coil mutuals first."""
    g = G['eig_passive']
    O, T = make(env, 'passive', g, vcoils=g['vcoils'])
    with pytest.raises(Exception, match='Coil mutuals required'):
        T.compute_Lmat()


@pytest.mark.parametrize('name', ['plate', 'cyl', 'torus', 'passive'])
def test_coil_sensor_operators_and_fr_goldens(env, name):
    g = G['fr_' + name]
    O, T = make(env, name, g, vcoils=g.get('vcoils'), icoils=g['icoils'])
    Mco = O.compute_Mcoil()
    Mcg = T.compute_Mcoil()
    assert relerr(Mcg, Mco) < 1e-12
    fl = [(ref_floop(R, Z), 1.0) for (R, Z) in g['floops']]
    Mso, Msco = O.compute_Msensor(fl)
    Msg, Mscg, _ = T.compute_Msensor(sensors=fl)
    assert relerr(Msg, Mso) < 1e-12 and relerr(Mscg, Msco) < 1e-12
    Lo = O.compute_Lmat()
    T.compute_Lmat()
    assert relerr(T.Lmat, Lo) < 1e-12
    T.compute_Rmat()
    R = T.Rmat.toarray()
    dc = 1.0 / MU0
    om = 2.0 * np.pi * g['freq']
    x = np.linalg.solve(1j * om * T.Lmat + R, 1j * (-om * Mcg[0] * dc))
    sig = np.stack([x.real, x.imag]) @ Msg
    sig[0] += dc * Mscg[0]
    assert np.abs(sig[0] / np.array(g['real']) - 1.0).max() < g['tol']
    assert np.abs(sig[1] / np.array(g['imag']) - 1.0).max() < g['tol']


def test_sensor_file_roundtrip(env, tmp_path):
    """compute_Msensor through the reference's floops.loc format (thin_wall.F90:2604-2618)."""
    from openfusiontoolkit_b200.ThinCurr.sensor import circular_flux_loop, save_sensors
    g = G['fr_plate']
    O, T = make(env, 'plate', g, icoils=g['icoils'])
    T.compute_Mcoil()
    sens = [circular_flux_loop(R, Z, 'FLOOP_%d' % k) for k, (R, Z) in enumerate(g['floops'])]
    path = str(tmp_path / 'floops.loc')
    save_sensors(sens, path)
    Ms, Msc, info = T.compute_Msensor(sensor_file=path)
    assert info['names'] == ['FLOOP_0', 'FLOOP_1']
    O.compute_Mcoil()
    Mso, Msco = O.compute_Msensor([(ref_floop(R, Z), 1.0) for (R, Z) in g['floops']])
    assert relerr(Ms, Mso) < 1e-12 and relerr(Msc, Msco) < 1e-12


def test_cross_coupling(env):
    """Mutual inductance between two models: plate (row) x passive dummy mesh moved close (column)."""
    from openfusiontoolkit_b200.ThinCurr import ThinCurr
    m1 = load_mesh('plate')
    r2, lc2 = dummy_mesh([0.1, 0.05, 0.08], size=0.6, nsplit=3)
    O1 = tw.OracleModel(m1['r'], m1['lc'], m1['reg'])
    O2 = tw.OracleModel(r2, lc2, None)
    Mo = O1.cross_coupling(O2)
    T1, T2 = ThinCurr(env), ThinCurr(env)
    T1.setup_model(r=m1['r'], lc=m1['lc'], reg=m1['reg'])
    T2.setup_model(r=r2, lc=lc2)
    Mg = T1.cross_coupling(T2)
    assert Mg.shape == (T1.nelems, T2.nelems) == Mo.shape
    # entry-wise 1e-10, plus the reference's own summation-order noise on entries that cancel:
    # two CPU summation orders of the oracle already differ by 6.6e-10 on an entry with |M| = 4e-8 A
    A = mutual_abs_sum(O1, O2)
    tol = 1e-10 * np.abs(Mo) + 64 * np.finfo(float).eps * A
    assert (np.abs(Mg - Mo) <= tol).all(), (np.abs(Mg - Mo) / tol).max()
    well = np.abs(Mo) > 1e-4 * A   # entries that lose < 4 digits to cancellation: plain relative error
    assert (np.abs(Mg - Mo)[well] / np.abs(Mo)[well]).max() < 1e-10
    assert relerr(Mg, Mo) < 1e-13
    # model x itself reproduces the self-inductance only up to the role rule of near pairs
    Ms = T1.cross_coupling(T1)
    T1.compute_Lmat()
    assert relerr(Ms, T1.Lmat) < 1e-4


@pytest.mark.parametrize('name', ['plate', 'torus', 'passive'])
def test_bmat(env, name):
    """B-field reconstruction operators.  The near field is a central finite difference of the
    analytic potential with step 1e-6 (thin_wall.F90:1995,2049-2075), which amplifies last-digit
    differences of log/atan2 by ~1e6: tolerance 1e-8 relative to the largest entry (SURVEY hard part 8)."""
    g = G['fr_' + name]
    O, T = make(env, name, g, vcoils=g.get('vcoils'), icoils=g['icoils'])
    if g.get('vcoils'):
        O.compute_Mcoil()
        T.compute_Mcoil()
    Bo, Bdo = O.compute_Bmat()
    Bg, Bdg = T.compute_Bmat()
    # reference memory layout Bel(nelems,np,3): Python view (3,nelems,np) over [comp][p][e] memory
    Bg_mem = np.asarray(Bg).reshape(3, T.np, T.nelems)
    assert relerr(Bg_mem, Bo) < 1e-8
    far = np.abs(Bo) < 1e-2 * np.abs(Bo).max()
    assert np.abs(Bg_mem - Bo)[far].max() / np.abs(Bo).max() < 1e-9
    Bdg_mem = np.asarray(Bdg).reshape(3, T.n_icoils, T.np)
    assert relerr(Bdg_mem, Bdo) < 1e-13


def test_bmat_cache_file(env, tmp_path):
    """compute_Bmat(cache_file): the HDF5 save file of thin_wall.F90:2208-2225 (MODEL_hash, Bel_X|Y|Z as [np][nelems],
    Bdr_X|Y|Z as [n_icoils][np]) is written, parses with the oracle-side HDF5 reader, and a second model loads it
    instead of rebuilding; a model with another mesh ignores it (hash mismatch)."""
    from oracle import h5min
    g = G['fr_torus']
    O, T = make(env, 'torus', g, vcoils=g.get('vcoils'), icoils=g['icoils'])
    fn = str(tmp_path / 'Bmat.save')
    Bg, Bdg = T.compute_Bmat(cache_file=fn)
    Bg = np.array(Bg).reshape(3, T.np, T.nelems)
    Bdg = np.array(Bdg).reshape(3, T.n_icoils, T.np)
    h = h5min.H5(fn)
    t = h.tree()
    assert sorted(t) == ['Bdr_X', 'Bdr_Y', 'Bdr_Z', 'Bel_X', 'Bel_Y', 'Bel_Z', 'MODEL_hash']
    mh = h.read(t['MODEL_hash'])
    assert mh.dtype == np.int32 and mh[0] == T.nelems and mh[1] == T.nc
    for c, nm in enumerate('XYZ'):
        assert np.array_equal(h.read(t['Bel_' + nm]), Bg[c])
        assert np.array_equal(h.read(t['Bdr_' + nm]), Bdg[c])
    O2, T2 = make(env, 'torus', g, vcoils=g.get('vcoils'), icoils=g['icoils'])
    B2, Bd2 = T2.compute_Bmat(cache_file=fn)       # loads
    assert np.array_equal(np.array(B2).reshape(Bg.shape), Bg) and np.array_equal(np.array(Bd2).reshape(Bdg.shape), Bdg)
    g3 = G['fr_plate']
    O3, T3 = make(env, 'plate', g3, vcoils=g3.get('vcoils'), icoils=g3['icoils'])
    B3, _ = T3.compute_Bmat(cache_file=fn)         # other model: hash mismatch => rebuilt (and the file rewritten)
    assert np.array(B3).size == 3 * T3.np * T3.nelems
    assert h5min.H5(fn).read(h5min.H5(fn).tree()['MODEL_hash'])[0] == T3.nelems
