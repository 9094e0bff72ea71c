"""ThinCurr CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/pure-Python restatement of the reference's model setup (tw_setup and the mesh
pre-processing it relies on) plus a ctypes front-end for the C restatement of the dense
operator builds in `thincurr_oracle.c`.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.

Reference: OpenFUSIONToolkit @ d08f001b (paths relative to /root/reference/src):
  physics/thin_wall.F90:166-523   tw_setup (holes, pmap, closures, qbasis)
  physics/thin_wall.F90:2230-2349 tw_setup_hole
  physics/thin_wall.F90:1690-1930 tw_compute_Rmat
  grid/mesh_local.F90:105-267,809-1092  edges / linkage / orientation sync / boundary
  grid/trimesh_type.F90:243-249,397-503,649-679  invert_cell / jacobian / norm / tang
Parity status: pinned against the reference's regression goldens (see
tests/test_oracle_golden.py); entry-wise operator values are not pinned by any
reference test.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libthincurr_oracle.so')
MU0 = np.pi * 4.e-7
TRI_ED = ((2, 1), (0, 2), (1, 0))  # trimesh_type.F90:34 tri_ed (0-based)


# CPU-baseline variants of the same source (bench.py only; the parity oracle is always the default build):
#   simd  : reference flags + the reference's `!$omp simd` loops (thin_wall.F90:1047,1070) enabled (-DTCO_SIMD)
#   tuned : -O3 -march=native on top of that ("what a tuned CPU build of the same loop nest reaches")
_VARIANTS = {None: ['-O2'], 'simd': ['-O2', '-DTCO_SIMD'], 'tuned': ['-O3', '-march=native', '-DTCO_SIMD']}


def build(force=False, variant=None):
    """Compile the C restatement with the reference's release flags (-O2 + OpenMP)."""
    src = os.path.join(_HERE, 'thincurr_oracle.c')
    hdr = os.path.join(_HERE, 'quad_tables.h')
    so = _SO if variant is None else _SO.replace('.so', '_%s.so' % variant)
    if (not force) and os.path.exists(so) and os.path.getmtime(so) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return so
    subprocess.check_call(['gcc'] + _VARIANTS[variant] + ['-fopenmp', '-fPIC', '-shared', '-std=c11', '-o', so, src, '-lm'])
    return so


def lmat_sample(model, i_begin, i_end, out, variant=None, nthreads=None):
    """bench.py: the tw_compute_LmatDirect loop over row cells [i_begin, i_end) with a CPU-baseline build variant and an
    explicit OpenMP thread count; returns the number of visited pairs."""
    L = ctypes.CDLL(build(variant=variant))
    L.tco_lmat_direct.restype = ctypes.c_longlong
    L.tco_lmat_direct.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    if nthreads:
        L.tco_set_num_threads(int(nthreads))
    return L.tco_lmat_direct(ctypes.byref(model.c), None, out.ctypes.data_as(ctypes.c_void_p), None, None, int(i_begin), int(i_end), 0, None)


class _CModel(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ('np', 'nc', 'np_active', 'nholes', 'n_vcoils', 'n_icoils', 'nelems', 'nfh')] + \
               [(n, ctypes.c_void_p) for n in ('r', 'lc', 'reg', 'ca', 'va', 'norm', 'qbasis', 'pmap', 'kfh', 'lfh', 'sens_mask')]


class _CCoils(ctypes.Structure):
    _fields_ = [('nsets', ctypes.c_int)] + \
               [(n, ctypes.c_void_p) for n in ('set_ptr', 'fil_ptr', 'pts', 'scales', 'radius', 'sens_mask')]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.tco_lmat_direct.restype = ctypes.c_longlong
        _lib.tco_lmat_direct.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.tco_phipot.restype = ctypes.c_double
        _lib.tco_pair_T.restype = ctypes.c_double
        _lib.tco_simple_hash.restype = ctypes.c_int32
        _lib.tco_simple_hash.argtypes = [ctypes.c_void_p, ctypes.c_long]
        _lib.tco_bel.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _lib.tco_filament_bfield.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong]
    return _lib


class CoilSets:
    """Flattened list of coil sets (each a list of polyline filaments)."""

    def __init__(self, sets):
        # sets: list of dict(filaments=[(pts[n,3], scale, radius, res_per_len)], sens_mask=bool)
        set_ptr, fil_ptr, pts, scales, radius, rpl, mask = [0], [0], [], [], [], [], []
        for s in sets:
            for (p, sc, rad, rp) in s['filaments']:
                p = np.asarray(p, dtype=np.float64).reshape(-1, 3)
                pts.append(p)
                fil_ptr.append(fil_ptr[-1] + len(p))
                scales.append(sc)
                radius.append(rad)
                rpl.append(rp)
            set_ptr.append(len(scales))
            mask.append(1 if s.get('sens_mask', False) else 0)
        self.nsets = len(sets)
        self.set_ptr = np.array(set_ptr, dtype=np.int32)
        self.fil_ptr = np.array(fil_ptr, dtype=np.int32)
        self.pts = np.ascontiguousarray(np.vstack(pts) if pts else np.zeros((0, 3)), dtype=np.float64)
        self.scales = np.array(scales, dtype=np.float64)
        self.radius = np.array(radius, dtype=np.float64)
        self.res_per_len = np.array(rpl, dtype=np.float64)
        self.sens_mask = np.array(mask, dtype=np.int32)
        self.c = _CCoils(self.nsets, self.set_ptr.ctypes.data, self.fil_ptr.ctypes.data, self.pts.ctypes.data,
                         self.scales.ctypes.data, self.radius.ctypes.data, self.sens_mask.ctypes.data)

    @staticmethod
    def concat(a, b):
        out = CoilSets([])
        out.nsets = a.nsets + b.nsets
        out.set_ptr = np.concatenate([a.set_ptr, b.set_ptr[1:] + a.set_ptr[-1]]).astype(np.int32)
        out.fil_ptr = np.concatenate([a.fil_ptr, b.fil_ptr[1:] + a.fil_ptr[-1]]).astype(np.int32)
        out.pts = np.ascontiguousarray(np.vstack([a.pts, b.pts]))
        out.scales = np.concatenate([a.scales, b.scales])
        out.radius = np.concatenate([a.radius, b.radius])
        out.res_per_len = np.concatenate([a.res_per_len, b.res_per_len])
        out.sens_mask = np.concatenate([a.sens_mask, b.sens_mask]).astype(np.int32)
        out.c = _CCoils(out.nsets, out.set_ptr.ctypes.data, out.fil_ptr.ctypes.data, out.pts.ctypes.data,
                        out.scales.ctypes.data, out.radius.ctypes.data, out.sens_mask.ctypes.data)
        return out


def circular_coil(R, Z, npts=181):
    """Default circular filament, thin_wall.F90:2499-2504."""
    k = np.arange(npts)
    theta = k * 2.0 * np.pi / float(npts - 1)
    return np.stack([R * np.cos(theta), R * np.sin(theta), Z * np.ones(npts)], 1)


class OracleModel:
    """Restatement of tw_type + tw_setup for linear triangle meshes."""

    def __init__(self, r, lc, reg=None, nodesets=(), closures=(), pmap=None, vcoils=None, icoils=None,
                 eta=None, sens_mask=None):
        r = np.ascontiguousarray(r, dtype=np.float64)
        if r.shape[1] == 2:
            r = np.hstack([r, np.zeros((len(r), 1))])
        self.r = np.ascontiguousarray(r)
        self.lc = np.array(lc, dtype=np.int32).copy()
        self.np_, self.nc = len(self.r), len(self.lc)
        self.reg = np.ones(self.nc, np.int32) if reg is None else np.asarray(reg, np.int32)
        self.nreg = int(self.reg.max())
        self.vcoils = vcoils if vcoils is not None else CoilSets([])
        self.icoils = icoils if icoils is not None else CoilSets([])
        # Icoil dummy radius removal, thin_wall.F90:202-206
        self.icoils.radius[:] = np.maximum(1.e-6, self.icoils.radius)
        self.n_vcoils, self.n_icoils = self.vcoils.nsets, self.icoils.nsets
        self.eta_surf = np.ones(self.nreg) if eta is None else np.asarray(eta, float) / MU0  # stored as eta/mu0 (:2864)
        self.sens_mask = np.zeros(self.nreg, np.int32) if sens_mask is None else np.asarray(sens_mask, np.int32)
        self._mesh_init()
        self._holes([np.asarray(n, np.int64) for n in nodesets])
        self._pmap(pmap, list(closures))
        self._geometry()
        self._cmodel()
        self.Ael2coil = self.Acoil2coil = self.Ael2dr = None

    # ---- bmesh_local_init(sync_normals=.TRUE.), mesh_local.F90:809-912
    def _mesh_init(self):
        lc, nc, npnt = self.lc, self.nc, self.np_
        # amesh_edges :105-205 -> edges numbered lexicographically by (lo,hi)
        ev = np.array([[lc[:, a], lc[:, b]] for (a, b) in TRI_ED])  # [slot,2,nc]
        lo = np.minimum(ev[:, 0], ev[:, 1]).T  # [nc,3]
        hi = np.maximum(ev[:, 0], ev[:, 1]).T
        key = lo.astype(np.int64) * npnt + hi
        uniq, inv = np.unique(key.ravel(), return_inverse=True)
        self.ne = len(uniq)
        self.le = np.stack([uniq // npnt, uniq % npnt], 1).astype(np.int64)
        self.lce = inv.reshape(nc, 3)  # |lce| (edge id per slot); sign not needed here
        # amesh_to_cell :213-267 -> point->cell and edge->cell in ascending cell order
        order = np.argsort(lc.ravel(), kind='stable')
        self.kpc = np.concatenate([[0], np.cumsum(np.bincount(lc.ravel(), minlength=npnt))])
        self.lpc = (order // 3).astype(np.int64)
        eorder = np.argsort(self.lce.ravel(), kind='stable')
        self.kec = np.concatenate([[0], np.cumsum(np.bincount(self.lce.ravel(), minlength=self.ne))])
        self.lec = (eorder // 3).astype(np.int64)
        # bmesh_neighbors :1022-1037
        lcc = -np.ones((nc, 3), np.int64)
        cnt = self.kec[1:] - self.kec[:-1]
        for c in range(nc):
            for j in range(3):
                e = self.lce[c, j]
                if cnt[e] == 2:
                    a, b = self.lec[self.kec[e]], self.lec[self.kec[e] + 1]
                    lcc[c, j] = a + b - c
        self.lcc = lcc
        self._sync_face_normals()
        # bmesh_boundary :1044-1092
        self.be = (cnt == 1)
        self.bp = np.zeros(npnt, bool)
        self.bp[self.le[self.be].ravel()] = True
        # amesh_interactions (point->edge part): edges of a point in ascending edge id
        pe = [[] for _ in range(npnt)]
        for e in range(self.ne):
            pe[self.le[e, 0]].append(e)
            pe[self.le[e, 1]].append(e)
        self.pe = pe
        self.edge_id = {(int(a), int(b)): e for e, (a, b) in enumerate(self.le)}

    def _invert(self, c):
        # trimesh_invert_cell, trimesh_type.F90:243-249
        self.lc[c, 1], self.lc[c, 2] = self.lc[c, 2], self.lc[c, 1]
        self.lce[c] = self.lce[c, [0, 2, 1]]
        self.lcc[c] = self.lcc[c, [0, 2, 1]]

    def _sync_face_normals(self):
        # mesh_local.F90:963-1014; the recursive DFS is unrolled with an explicit stack
        # that preserves the visiting order (slot 1..3, depth first).
        oriented = np.zeros(self.nc, bool)
        self.nflipped = 0
        for seed in range(self.nc):
            if oriented[seed]:
                continue
            oriented[seed] = True
            stack = [[seed, 0]]
            while stack:
                f1, j = stack[-1]
                if j == 3:
                    stack.pop()
                    continue
                stack[-1][1] += 1
                f2 = self.lcc[f1, j]
                if f2 < 0 or oriented[f2]:
                    continue
                ed1 = (self.lc[f1, TRI_ED[j][0]], self.lc[f1, TRI_ED[j][1]])
                k = [kk for kk in range(3) if self.lcc[f2, kk] == f1][0]
                ed2 = (self.lc[f2, TRI_ED[k][0]], self.lc[f2, TRI_ED[k][1]])
                if ed1 == ed2:
                    self._invert(f2)
                    self.nflipped += 1
                oriented[f2] = True
                stack.append([f2, 0])

    # ---- geometry: trimesh_jacobian/norm/tang + qbasis (thin_wall.F90:343-352)
    def _geometry(self):
        P = self.r[self.lc]  # nc,3,3
        t1 = P[:, 1] - P[:, 0]
        t1 = t1 / np.sqrt((t1 ** 2).sum(1))[:, None]
        t2 = P[:, 2] - P[:, 0]
        t2 = t2 - (t2 * t1).sum(1)[:, None] * t1
        t2 = t2 / np.sqrt((t2 ** 2).sum(1))[:, None]
        nrm = np.cross(t1, t2)
        proj = np.stack([(P * t1[:, None, :]).sum(-1), (P * t2[:, None, :]).sum(-1)], -1)  # nc,3,2
        A = np.stack([proj[:, 1] - proj[:, 0], proj[:, 2] - proj[:, 0]], 1)
        det = A[:, 0, 0] * A[:, 1, 1] - A[:, 0, 1] * A[:, 1, 0]
        C = np.empty_like(A)
        C[:, 0, 0] = A[:, 1, 1]
        C[:, 1, 1] = A[:, 0, 0]
        C[:, 0, 1] = -A[:, 0, 1]
        C[:, 1, 0] = -A[:, 1, 0]
        C = C / det[:, None, None]
        g2 = C[:, 0, 0, None] * t1 + C[:, 1, 0, None] * t2
        g3 = C[:, 0, 1, None] * t1 + C[:, 1, 1, None] * t2
        g1 = -(g2 + g3)
        gop = np.stack([g1, g2, g3], 1)
        self.ca = np.abs(det / 2.0)
        self.norm = np.ascontiguousarray(nrm)
        self.qbasis = np.ascontiguousarray(np.cross(gop, nrm[:, None, :]))  # [nc][vert][xyz]
        va = np.zeros(self.np_)
        for k in range(3):
            np.add.at(va, self.lc[:, k], self.ca / 3.0)
        self.va = va

    def _cell_norm(self, c):
        P = self.r[self.lc[c]]
        t1 = P[1] - P[0]
        t1 = t1 / np.sqrt((t1 ** 2).sum())
        t2 = P[2] - P[0]
        t2 = t2 - (t2 @ t1) * t1
        t2 = t2 / np.sqrt((t2 ** 2).sum())
        return np.cross(t1, t2)

    def _findedge(self, a, b):
        return self.edge_id.get((min(a, b), max(a, b)), -1)

    # ---- holes, thin_wall.F90:207-281, 405-522, 2230-2349
    def _hole_pseq(self, i0):
        if not self.bp[i0]:
            raise RuntimeError('Hole starting vertex is not on boundary')
        ipt, chain, eprev = i0, [i0], -1
        for _ in range(int(self.be.sum())):
            for ed in self.pe[ipt]:
                if ed == eprev or not self.be[ed]:
                    continue
                ipt = int(self.le[ed].sum() - ipt)
                chain.append(ipt)
                eprev = ed
                break
            if ipt == i0:
                break
        if ipt != i0:
            raise RuntimeError('could not find periodic path')
        return chain[:-1]

    def _order_hole_list(self, list_in):
        n = len(list_in)
        srt = sorted(int(v) for v in list_in)
        pos = {v: i for i, v in enumerate(srt)}
        flag = [0] * n
        flag[0] = 1
        ipt = srt[0]
        out = [ipt]
        eprev = -1
        for jj in range(1, n + 1):
            if jj == n - 2:
                flag[0] = 0
            last_item = None
            for ed in self.pe[ipt]:
                if ed == eprev:
                    continue
                ptp = int(self.le[ed].sum() - ipt)
                cand = pos.get(ptp)
                if cand is None or flag[cand] == 1:
                    continue
                nlinks = 0
                for ed2 in self.pe[ptp]:
                    ptp2 = int(self.le[ed2].sum() - ptp)
                    c2 = pos.get(ptp2)
                    if c2 is None or flag[c2] == 1:
                        continue
                    nlinks += 1
                last_item = (ptp, cand, ed)
                if nlinks > 1:
                    continue
                last_item = None
                flag[cand] = 1
                ipt = ptp
                if jj < n:
                    out.append(ipt)
                    eprev = ed
                break
            if last_item is not None:
                flag[last_item[1]] = 1
                ipt = last_item[0]
                if jj < n:
                    out.append(ipt)
                    eprev = last_item[2]
        if ipt != srt[0]:
            raise RuntimeError('hole path is not periodic')
        return out

    def _setup_hole(self, lp):
        n = len(lp)
        fo = np.zeros(self.nc, np.int64)
        po = np.zeros(n, np.int64)
        for i in range(n):
            a, b = lp[i], lp[(i + 1) % n]
            k = self._findedge(a, b)
            if k < 0:
                raise RuntimeError('Could not find edge')
            evec = self.r[b] - self.r[a]
            ecc = (self.r[b] + self.r[a]) / 2.0
            cells = self.lec[self.kec[k]:self.kec[k + 1]]
            for c in cells:
                if fo[c] != 0:
                    continue
                P = self.r[self.lc[c]]
                ptcc = (P[0] + P[1] + P[2]) / 3.0
                val = np.cross(ptcc - ecc, evec) @ self._cell_norm(c)
                fo[c] = 1 if (val >= 0.0 and not (val == 0.0 and np.signbit(val))) else -1
            if self.be[k]:
                po[i] = fo[cells[0]]
                po[(i + 1) % n] = fo[cells[0]]
        if (po >= 0).all():
            po[:] = 1
        elif (po < 0).all():
            po[:] = -1
        else:
            prev = 0
            for i in range(n):
                if po[i] == 0:
                    if prev != 0:
                        po[i] = prev
                else:
                    prev = po[i]
            for i in range(n):
                if po[i] != 0:
                    break
                po[i] = prev
        for _sweep in range(10):
            ok = True
            for v in lp:
                for c in self.lpc[self.kpc[v]:self.kpc[v + 1]]:
                    if fo[c] == 0:
                        for l in range(3):
                            f = self.lcc[c, l]
                            if f < 0:
                                continue
                            if fo[f] != 0:
                                fo[c] = fo[f]
                                break
                        if fo[c] == 0:
                            ok = False
            if ok:
                break
        if not ok:
            raise RuntimeError('Error orienting cells')
        out = []  # (chain position, cell, sign)
        for i, v in enumerate(lp):
            for c in self.lpc[self.kpc[v]:self.kpc[v + 1]]:
                if fo[c] == po[i]:
                    out.append((i, int(c), int(fo[c])))
        return out

    def _holes(self, nodesets):
        self.nholes = len(nodesets)
        self.hole_chains, self.hole_cells = [], []
        per_cell = [[] for _ in range(self.nc)]
        for h, ns in enumerate(nodesets):
            lp = self._hole_pseq(int(ns[0])) if len(ns) == 1 else self._order_hole_list(ns)
            self.hole_chains.append(lp)
            cells = self._setup_hole(lp)
            self.hole_cells.append(cells)
            for (i, c, sg) in cells:
                l = [k for k in range(3) if self.lc[c, k] == lp[i]][0]
                per_cell[c].append((sg * (h + 1), l))
        kfh = [0]
        lfh = []
        for c in range(self.nc):
            lfh += per_cell[c]
            kfh.append(len(lfh))
        self.kfh = np.array(kfh, np.int32)
        self.lfh = np.array(lfh, np.int32).reshape(-1, 2)
        self.nfh = len(lfh)

    # ---- DOF map, thin_wall.F90:282-320
    def _pmap(self, pmap, closures):
        if pmap is None:
            pm = np.zeros(self.np_, np.int64)
            act = ~self.bp
            pm[act] = np.arange(1, act.sum() + 1)
            self.closure_verts = []
            for ci in closures:
                l, j = -1, 0
                for k in range(3):
                    v = self.lc[ci, k]
                    if pm[v] <= 0:
                        continue
                    cnt = self.kpc[v + 1] - self.kpc[v]
                    if cnt > l:
                        l, j = cnt, k
                v = self.lc[ci, j]
                if pm[v] == 0:
                    raise RuntimeError('Error getting closure vertex')
                pm[v] = -1
                self.closure_verts.append(int(v))
            act = pm > 0
            pm[:] = 0
            pm[act] = np.arange(1, act.sum() + 1)
            self.np_active = int(act.sum())
        else:
            pm = np.asarray(pmap, np.int64).copy()
            self.np_active = int(pm.max())
        self.pmap = pm.astype(np.int32)
        self.nelems = self.np_active + self.nholes + self.n_vcoils

    def _cmodel(self):
        self.lc = np.ascontiguousarray(self.lc, np.int32)
        self.reg = np.ascontiguousarray(self.reg, np.int32)
        self.c = _CModel(self.np_, self.nc, self.np_active, self.nholes, self.n_vcoils, self.n_icoils, self.nelems,
                         self.nfh, self.r.ctypes.data, self.lc.ctypes.data, self.reg.ctypes.data, self.ca.ctypes.data,
                         self.va.ctypes.data, self.norm.ctypes.data, self.qbasis.ctypes.data, self.pmap.ctypes.data,
                         self.kfh.ctypes.data, self.lfh.ctypes.data if self.nfh else None, self.sens_mask.ctypes.data)

    # ---- operators -----------------------------------------------------------------
    def compute_Mcoil(self):
        """tw_compute_Ael2dr + tw_compute_Lmat_coils (thin_wall.F90:567-883).
        Returns Ael2dr viewed as Python does: (n_icoils, nelems)."""
        L = lib()
        allc = CoilSets.concat(self.vcoils, self.icoils)
        ntot = allc.nsets
        tmp = np.zeros((max(ntot, 1), self.nelems))  # Fortran (nelems, ntot)
        self.nrad_cross = np.zeros(max(ntot, 1), np.int32)
        if ntot:
            L.tco_ael2coil(ctypes.byref(self.c), ctypes.byref(allc.c), tmp.ctypes.data_as(ctypes.c_void_p),
                           self.nrad_cross.ctypes.data_as(ctypes.c_void_p))
        self.Ael2coil = np.ascontiguousarray(tmp[:self.n_vcoils])          # [v][e]  == Fortran (nelems,n_v)
        Ael2dr = np.ascontiguousarray(tmp[self.n_vcoils:ntot])             # [i][e]
        A = np.zeros((max(ntot, 1), max(self.n_vcoils, 1)))                # Fortran (n_v, ntot): A[j][l]
        if self.n_vcoils and ntot:
            L.tco_filament_mutual(ctypes.byref(self.vcoils.c), ctypes.byref(allc.c), 1, A.ctypes.data_as(ctypes.c_void_p))
        self.Acoil2coil = np.ascontiguousarray(A[:self.n_vcoils, :self.n_vcoils])  # [j][l] = Acoil2coil(l,j)
        self.vcoil_Lself = np.array([A[i, i] for i in range(self.n_vcoils)])
        ns = self.np_active + self.nholes
        for i in range(self.n_vcoils):
            for jj in range(self.n_icoils):
                Ael2dr[jj, ns + i] = A[jj + self.n_vcoils, i]
        self.Ael2dr = Ael2dr * MU0 / (4.0 * np.pi)
        return self.Ael2dr

    def compute_Lmat(self, i_begin=0, i_end=None, finalize=True, hist=None, out=None):
        """tw_compute_LmatDirect self-inductance (thin_wall.F90:887-1186)."""
        if self.n_vcoils > 0 and self.Acoil2coil is None:
            raise RuntimeError('Coil mutuals required if, # of Vcoils > 0')
        N = self.nelems
        Lm = np.zeros((N, N)) if out is None else out
        a2c = self.Ael2coil.ctypes.data if self.n_vcoils else None
        c2c = self.Acoil2coil.ctypes.data if self.n_vcoils else None
        i_end = self.nc if i_end is None else i_end
        hp = hist.ctypes.data_as(ctypes.c_void_p) if hist is not None else None
        self.visited = lib().tco_lmat_direct(ctypes.byref(self.c), None, Lm.ctypes.data_as(ctypes.c_void_p), a2c, c2c,
                                             int(i_begin), int(i_end), 1 if finalize else 0, hp)
        if finalize and out is None:
            self.Lmat = Lm
        return Lm

    def lmat_rows(self, dofs, with_abs=False):
        """Rows of the self-inductance matrix (vertex/hole DOFs, 0-based) from the per-entry
        definition; independent of the loop-nest restatement in compute_Lmat.
        with_abs: also the sum of the magnitudes of the terms of every entry (conditioning of the sum)."""
        dofs = np.ascontiguousarray(dofs, np.int32)
        out = np.zeros((len(dofs), self.nelems))
        if not with_abs:
            lib().tco_lmat_rows(ctypes.byref(self.c), ctypes.c_int(len(dofs)), dofs.ctypes.data_as(ctypes.c_void_p),
                                out.ctypes.data_as(ctypes.c_void_p))
            return out
        ab = np.zeros_like(out)
        lib().tco_lmat_rows2(ctypes.byref(self.c), ctypes.c_int(len(dofs)), dofs.ctypes.data_as(ctypes.c_void_p),
                             out.ctypes.data_as(ctypes.c_void_p), ab.ctypes.data_as(ctypes.c_void_p))
        return out, ab

    def cross_coupling(self, other):
        """tw_compute_LmatDirect(self, M, col_model=other); returns Python view (self.nelems, other.nelems)."""
        M = np.zeros((self.nelems, other.nelems))  # Fortran (other.nelems, self.nelems)
        lib().tco_lmat_direct(ctypes.byref(self.c), ctypes.byref(other.c), M.ctypes.data_as(ctypes.c_void_p), None, None,
                              0, self.nc, 1, None)
        return M

    def compute_Msensor(self, floops):
        """tw_compute_mutuals (thin_wall.F90:1418-1686).  floops: list of (pts[n,3], scale_fac).
        Returns (Ael2sen viewed (nelems,nsens), Adr2sen viewed (n_icoils,nsens))."""
        L = lib()
        sens = CoilSets([dict(filaments=[(p, sf, 0.0, 0.0)]) for (p, sf) in floops])
        ns = sens.nsets
        Ael2sen = np.zeros((self.nelems, max(ns, 1)))  # Fortran (nsens, nelems)
        if ns:
            L.tco_ael2sen(ctypes.byref(self.c), ctypes.byref(sens.c), Ael2sen.ctypes.data_as(ctypes.c_void_p))
        allc = CoilSets.concat(self.vcoils, self.icoils)
        ntot = allc.nsets
        A = np.zeros((max(ntot, 1), max(ns, 1)))  # Fortran (nsens, ntot): A[j][i]
        if ns and ntot:
            L.tco_filament_mutual(ctypes.byref(sens.c), ctypes.byref(allc.c), 0, A.ctypes.data_as(ctypes.c_void_p))
            for j in range(ns):
                A[:, j] *= sens.scales[j]
        Adr2sen = A[self.n_vcoils:ntot].copy() * MU0 / (4.0 * np.pi)
        nsd = self.np_active + self.nholes
        if ns:
            for i in range(self.n_vcoils):
                Ael2sen[nsd + i, :] = A[i, :]
            Ael2sen = Ael2sen / (4.0 * np.pi)
        return Ael2sen[:, :ns], Adr2sen[:, :ns]

    def compute_Bmat(self, i_begin=0, i_end=None, finalize=True):
        """tw_compute_Bops (thin_wall.F90:1989-2169).  Returns (Bel (3,np,nelems) memory order
        [comp][p][e] == Fortran (nelems,np,3); Bdr (3,n_icoils,np) == Fortran (np,n_icoils,3))."""
        L = lib()
        N, npnt = self.nelems, self.np_
        Bel = np.zeros((3, npnt, N))
        i_end = self.nc if i_end is None else i_end
        L.tco_bel(ctypes.byref(self.c), Bel.ctypes.data_as(ctypes.c_void_p), int(i_begin), int(i_end))
        if not finalize:
            return Bel, None
        if self.n_vcoils:
            off = (self.np_active + self.nholes) * 8
            L.tco_filament_bfield(ctypes.byref(self.c), ctypes.byref(self.vcoils.c),
                                  ctypes.c_void_p(Bel.ctypes.data + off), N, 1, npnt * N)
        Bel /= (4.0 * np.pi)
        Bdr = np.zeros((3, max(self.n_icoils, 1), npnt))
        if self.n_icoils:
            L.tco_filament_bfield(ctypes.byref(self.c), ctypes.byref(self.icoils.c), Bdr.ctypes.data_as(ctypes.c_void_p),
                                  1, npnt, npnt * self.n_icoils)
        Bdr = Bdr[:, :self.n_icoils] * MU0 / (4.0 * np.pi)
        return Bel, Bdr

    # ---- SURVEY 8f rows ---------------------------------------------------------------
    def cross_eval(self, other, field, counts=None):
        """tw_compute_Lmat_MF(self, other, nrhs, a, b) as ThinCurr.cross_eval sees it (thin_wall.F90:1190-1414,
        _core.py:533-549): field (nrhs, self.nelems) -> (nrhs, other.nelems)."""
        a = np.ascontiguousarray(field, np.float64)
        b = np.zeros((a.shape[0], other.nelems))
        cp = counts.ctypes.data_as(ctypes.c_void_p) if counts is not None else None
        lib().tco_lmat_mf(ctypes.byref(self.c), ctypes.byref(other.c), ctypes.c_int(a.shape[0]), a.ctypes.data_as(ctypes.c_void_p),
                          b.ctypes.data_as(ctypes.c_void_p), cp)
        return b

    def block(self, pts):
        """oft_tw_block of a vertex subset (thin_wall_hodlr.F90:60-70): 0-based vertex ids -> (cells touching them in
        ascending order, inv_map [np] with the 1-based position in the block or 0)."""
        pts = np.asarray(pts, np.int32)
        inv = np.zeros(self.np_, np.int32)
        inv[pts] = np.arange(1, len(pts) + 1)
        cells = np.nonzero((inv[self.lc] > 0).any(axis=1))[0].astype(np.int32)
        return np.ascontiguousarray(cells), inv

    def lmat_block(self, row_pts, col_pts, other=None):
        """tw_compute_Lmatblock (thin_wall_hodlr.F90:289-404); returns [len(row_pts)][len(col_pts)]
        (= Fortran Lmat(col, row))."""
        col = self if other is None else other
        rc, rinv = self.block(row_pts)
        cc, cinv = col.block(col_pts)
        out = np.zeros((len(row_pts), len(col_pts)))
        lib().tco_lmat_block(ctypes.byref(self.c), ctypes.byref(col.c), ctypes.c_int(len(rc)), rc.ctypes.data_as(ctypes.c_void_p),
                             rinv.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(cc)), cc.ctypes.data_as(ctypes.c_void_p),
                             cinv.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(row_pts)), ctypes.c_int(len(col_pts)),
                             out.ctypes.data_as(ctypes.c_void_p))
        return out

    def lmat_hole(self):
        """tw_compute_LmatHole(self, self) (thin_wall_hodlr.F90:136-285); returns [nholes + n_vcoils][nelems]."""
        out = np.zeros((self.nholes + self.n_vcoils, self.nelems))
        a2c = self.Ael2coil.ctypes.data_as(ctypes.c_void_p) if self.n_vcoils else None
        c2c = self.Acoil2coil.ctypes.data_as(ctypes.c_void_p) if self.n_vcoils else None
        lib().tco_lmat_hole(ctypes.byref(self.c), ctypes.byref(self.c), a2c, c2c, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def bops_block(self, row_pts, col_pts, direction):
        """tw_compute_Bops_block (thin_wall_hodlr.F90:580-691), direction 0/1/2; returns [len(row_pts)][len(col_pts)]."""
        rc, rinv = self.block(row_pts)
        cp = np.ascontiguousarray(col_pts, np.int32)
        out = np.zeros((len(row_pts), len(cp)))
        lib().tco_bops_block(ctypes.byref(self.c), ctypes.c_int(len(rc)), rc.ctypes.data_as(ctypes.c_void_p),
                             rinv.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(row_pts)), ctypes.c_int(len(cp)),
                             cp.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(direction), out.ctypes.data_as(ctypes.c_void_p))
        return out

    def cell_basis(self):
        """Per-cell list of (dof0, E vector): vertex DOFs + signed hole DOFs (A.3 of SURVEY)."""
        out = []
        for c in range(self.nc):
            d = {}
            for l in range(3):
                k = self.pmap[self.lc[c, l]]
                if k > 0:
                    d[k - 1] = d.get(k - 1, 0) + self.qbasis[c, l]
            for ii in range(self.kfh[c], self.kfh[c + 1]):
                h, l = self.lfh[ii]
                k = self.np_active + abs(h) - 1
                d[k] = d.get(k, 0) + np.sign(h) * self.qbasis[c, l]
            out.append(d)
        return out

    def compute_Rmat(self):
        """tw_compute_Rmat values (thin_wall.F90:1868-1911) as a scipy CSR matrix."""
        import scipy.sparse as sp
        rows, cols, vals = [], [], []
        for c, d in enumerate(self.cell_basis()):
            eta = self.eta_surf[self.reg[c] - 1]
            ks = list(d)
            for a in ks:
                for b in ks:
                    rows.append(a)
                    cols.append(b)
                    vals.append(eta * (d[a] @ d[b]) * self.ca[c])
        ns = self.np_active + self.nholes
        for i in range(self.n_vcoils):
            Rself = 0.0
            for k in range(self.vcoils.set_ptr[i], self.vcoils.set_ptr[i + 1]):
                p = self.vcoils.pts[self.vcoils.fil_ptr[k]:self.vcoils.fil_ptr[k + 1]]
                dl = np.sqrt(((p[1:] - p[:-1]) ** 2).sum(1)).sum()
                Rself += self.vcoils.res_per_len[k] * dl
            rows.append(ns + i)
            cols.append(ns + i)
            vals.append(Rself / MU0)
        self.Rmat = sp.csr_array(sp.coo_array((vals, (rows, cols)), shape=(self.nelems, self.nelems)))
        return self.Rmat

    def get_eigs(self, neigs):
        """Leading L x = tau R x modes (thin_wall_solvers.F90:39-115, sorted by |tau|)."""
        import scipy.linalg as sl
        w = sl.eigh(self.Lmat, self.Rmat.toarray(), eigvals_only=True)
        w = w[np.argsort(-np.abs(w))]
        return w[:neigs]

    def hash_lc(self):
        lc1 = np.ascontiguousarray(self.lc + 1, np.int32)
        return lib().tco_simple_hash(lc1.ctypes.data, lc1.nbytes)

    def hash_r(self):
        return lib().tco_simple_hash(self.r.ctypes.data, self.r.nbytes)
