"""`ThinCurr` -- same class name, method names, argument meaning, return shapes and error
behaviour as the reference's src/python/OpenFUSIONToolkit/ThinCurr/_core.py for the operator
build path (setup_model :99-166, compute_Lmat :269-288, compute_Bmat :307-331, compute_Mcoil
:333-348, compute_Msensor :350-387, compute_Rmat :491-513, cross_coupling :515-531).  All
matrices are zero-copy numpy views of library-owned buffers in the reference's layouts.
`get_eigs` (iterative path, _core.py:563-580) and `apply_Lmat` (:472-489 region) run on the device-resident matrix.
`cross_eval` (:533-549, matrix-free mutual apply) and `build_reduced_model` (:717-739) run on the device as well.
Methods of the reference that belong to the remaining downstream solvers (time stepping, frequency response, plotting)
are out of scope and raise NotImplementedError.
"""
import ctypes
from ctypes import c_bool, c_double, c_int, c_void_p

import numpy
import scipy.sparse

from .._interface import (B200_APPLY_FN, b200_device_alloc, b200_device_free, b200_ipc_close, b200_ipc_export, b200_ipc_open,
                          b200_Lmat_exchange, b200_Lmat_gather, b200_Lmat_save_begin, b200_Lmat_save_rows, b200_lr_eigs,
                          b200_release_device, b200_rows_apply, b200_rows_to_host, thincurr_apply_Lmat, thincurr_eigenvalues,
                          b200_Lmat_block, b200_Bel_shard, b200_Lmat_shard, b200_Lmat_shard_host, b200_Lmat_shard_sym, b200_shard_rows_sym, b200_destroy, b200_get_model,
                          b200_hashes, b200_last_error, b200_msensor, b200_pair_stats, b200_plan, b200_set_coils,
                          b200_set_sensors, b200_setup, b200_shard_rows, c_double_ptr, c_int_ptr, mu0, oftpy_load_xml,
                          thincurr_Bmat, thincurr_cross_coupling, thincurr_get_eta, thincurr_get_sensor_name,
                          thincurr_Lmat, thincurr_Mcoil, thincurr_Msensor, thincurr_Rmat, thincurr_set_eta,
                          thincurr_setup, thincurr_cross_eval, thincurr_reduce_model)


def _check(rc):
    if rc != 0:
        raise Exception(b200_last_error().decode())


class ThinCurr():
    '''! ThinCurr thin-wall E-M model class (B200 operator-build backend)'''

    def __init__(self, OFT_env):
        self._oft_env = OFT_env
        self.tw_obj = c_void_p()
        self.nregs = -1
        self.np = -1
        self.ne = -1
        self.nc = -1
        self.np_active = -1
        self.nholes = -1
        self.n_vcoils = -1
        self.nelems = -1
        self.n_icoils = -1
        self.Lmat = None
        self.Lmat_hodlr = c_void_p()
        self.Rmat = None
        self.r = None
        self.lc = None
        self.reg = None
        self._xml_ptr = c_void_p()
        self._sensor_ptr = c_void_p()

    def __del__(self):
        try:
            if self.tw_obj:
                b200_destroy(self.tw_obj)
                self.tw_obj = c_void_p()
        except Exception:
            pass

    def _set_sizes(self, sizes):
        (self.np, self.ne, self.nc, self.nregs, self.np_active, self.nholes, self.n_vcoils, self.nelems,
         self.n_icoils) = [int(v) for v in sizes]

    def setup_model(self, r=None, lc=None, reg=None, mesh_file=None, pmap=None, xml_filename=None, jumper_start=0,
                    nodesets=None, closures=None):
        '''! Setup ThinCurr model (reference: _core.py:99-166).

        @param r Point list `(np,3)`
        @param lc Cell list `(nc,3)` (0-based)
        @param reg Region tag `(nc,)`
        @param mesh_file File containing model in native mesh format
        @param pmap Point map for periodic grids
        @param xml_filename Path to XML file for model
        @param jumper_start Index of first jumper nodeset in meshfile
        @param nodesets (extension) hole nodesets as lists of 0-based vertex ids for in-memory meshes
        @param closures (extension) closure cells (0-based) for in-memory closed meshes
        '''
        if self.nregs != -1:
            raise ValueError('Mesh already setup, delete or create new instance for new model')
        if xml_filename is not None:
            oftpy_load_xml(self._oft_env.path2c(xml_filename), ctypes.byref(self._xml_ptr))
        sizes = numpy.zeros((9,), dtype=numpy.int32)
        if mesh_file is not None:
            if (r is not None) or (lc is not None) or (reg is not None):
                raise ValueError('Specification of "mesh_file" is incompatible with specification of "r", "lc", and "reg"')
            rfake = numpy.ones((1, 1), dtype=numpy.float64)
            lcfake = numpy.ones((1, 1), dtype=numpy.int32)
            regfake = numpy.ones((1,), dtype=numpy.int32)
            pmap = -numpy.ones((1,), dtype=numpy.int32) if pmap is None else numpy.ascontiguousarray(pmap, dtype=numpy.int32)
            error_string = self._oft_env.get_c_errorbuff()
            thincurr_setup(self._oft_env.path2c(mesh_file), c_int(-1), rfake, c_int(-1), lcfake, regfake, pmap,
                           c_int(jumper_start), ctypes.byref(self.tw_obj), sizes, error_string, self._xml_ptr)
            if error_string.value != b'':
                raise Exception(error_string.value.decode())
        elif r is not None:
            if lc is None:
                raise ValueError('"r" and "lc" must be both be specified')
            if jumper_start != 0:
                raise ValueError('"jumper_start" not supported with manual mesh specification')
            r = numpy.ascontiguousarray(r, dtype=numpy.float64)
            if r.shape[1] == 2:
                r = numpy.ascontiguousarray(numpy.hstack([r, numpy.zeros((r.shape[0], 1))]))
            lc = numpy.ascontiguousarray(lc, dtype=numpy.int32)
            reg = numpy.ones((lc.shape[0],), dtype=numpy.int32) if reg is None else numpy.ascontiguousarray(reg, dtype=numpy.int32)
            if nodesets is None and closures is None and pmap is None:
                error_string = self._oft_env.get_c_errorbuff()
                thincurr_setup(self._oft_env.path2c(''), c_int(r.shape[0]), r, c_int(lc.shape[0]), lc + 1, reg,
                               -numpy.ones((1,), dtype=numpy.int32), c_int(0), ctypes.byref(self.tw_obj), sizes,
                               error_string, self._xml_ptr)
                if error_string.value != b'':
                    raise Exception(error_string.value.decode())
            else:
                nodesets = [] if nodesets is None else nodesets
                ptr = numpy.zeros(len(nodesets) + 1, dtype=numpy.int32)
                for k, ns in enumerate(nodesets):
                    ptr[k + 1] = ptr[k] + len(ns)
                val = numpy.ascontiguousarray(numpy.concatenate([numpy.asarray(ns).ravel() for ns in nodesets]) + 1
                                              if len(nodesets) else numpy.zeros(1), dtype=numpy.int32)
                cl = numpy.ascontiguousarray((numpy.asarray(closures).ravel() + 1) if closures is not None and len(closures)
                                             else numpy.zeros(1), dtype=numpy.int32)
                ncl = 0 if closures is None else len(closures)
                pm = None if pmap is None else numpy.ascontiguousarray(pmap, dtype=numpy.int32)
                _check(b200_setup(r.shape[0], r, lc.shape[0], numpy.ascontiguousarray(lc + 1), reg.ctypes.data_as(c_void_p),
                                  pm.ctypes.data_as(c_void_p) if pm is not None else None, len(nodesets), ptr, val, ncl, cl,
                                  self._xml_ptr, ctypes.byref(self.tw_obj), sizes))
        else:
            raise ValueError('Mesh filename (native format) or mesh values (r, lc) required')
        self._set_sizes(sizes)

    # ---- extensions for in-memory coil / sensor definitions ------------------------------------
    def set_coils(self, kind, coil_sets):
        '''! (extension) Define V-coils (`kind='vcoil'`) or I-coils (`'icoil'`) from memory.

        @param coil_sets list of sets; a set is a list of dicts(pts[n,3], scale, radius, res_per_len)
        '''
        set_ptr, fil_ptr, pts, sc, rad, rpl = [0], [0], [], [], [], []
        for s in coil_sets:
            for f in s:
                p = numpy.asarray(f['pts'], dtype=numpy.float64).reshape(-1, 3)
                pts.append(p)
                fil_ptr.append(fil_ptr[-1] + p.shape[0])
                sc.append(f.get('scale', 1.0))
                rad.append(f.get('radius', -1.0))
                rpl.append(f.get('res_per_len', -1.0))
            set_ptr.append(len(sc))
        sizes = numpy.zeros((9,), dtype=numpy.int32)
        P = numpy.ascontiguousarray(numpy.vstack(pts) if pts else numpy.zeros((1, 3)))
        _check(b200_set_coils(self.tw_obj, 0 if kind == 'vcoil' else 1, len(coil_sets), numpy.array(set_ptr, dtype=numpy.int32),
                              numpy.array(fil_ptr, dtype=numpy.int32), P, numpy.array(sc + [0.0], dtype=numpy.float64),
                              numpy.array(rad + [0.0], dtype=numpy.float64), numpy.array(rpl + [0.0], dtype=numpy.float64),
                              numpy.zeros(len(coil_sets) + 1, dtype=numpy.int32), sizes))
        self._set_sizes(sizes)

    def compute_Lmat(self, cache_file=None, use_hodlr=False):
        '''! Compute the self-inductance matrix for this model (reference: _core.py:269-288)'''
        cache_string = self._oft_env.path2c("" if cache_file is None else cache_file)
        Lmat_loc = c_void_p()
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_Lmat(self.tw_obj, use_hodlr, ctypes.byref(Lmat_loc), cache_string, error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())
        self.Lmat = numpy.ctypeslib.as_array(ctypes.cast(Lmat_loc, c_double_ptr), shape=(self.nelems, self.nelems))

    def compute_Bmat(self, cache_file=None):
        '''! Magnetic field reconstruction operators (reference: _core.py:307-331).  Returned views have
        the reference's shapes `(3,nelems,np)` / `(3,n_icoils,np)` over Fortran `Bel(nelems,np,3)` /
        `Bdr(np,n_icoils,3)` memory.'''
        cache_string = self._oft_env.path2c("" if cache_file is None else cache_file)
        Bmat_loc = c_void_p()
        Bdr_ptr = c_void_p()
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_Bmat(self.tw_obj, c_void_p(), ctypes.byref(Bmat_loc), ctypes.byref(Bdr_ptr), cache_string, error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())
        return numpy.ctypeslib.as_array(ctypes.cast(Bmat_loc, c_double_ptr), shape=(3, self.nelems, self.np)), \
            numpy.ctypeslib.as_array(ctypes.cast(Bdr_ptr, c_double_ptr), shape=(3, self.n_icoils, self.np))

    def compute_Mcoil(self, cache_file=None):
        '''! Mutual inductance between passive (mesh+Vcoils) and active elements (Icoils) `(n_icoils,nelems)`'''
        cache_string = self._oft_env.path2c("" if cache_file is None else cache_file)
        Mc_loc = c_void_p()
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_Mcoil(self.tw_obj, ctypes.byref(Mc_loc), cache_string, error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())
        return numpy.ctypeslib.as_array(ctypes.cast(Mc_loc, c_double_ptr), shape=(self.n_icoils, self.nelems))

    def compute_Msensor(self, sensor_file=None, cache_file=None, sensors=None):
        '''! Mutual inductance between model and sensors (reference: _core.py:350-387).
        `sensors` (extension): list of (pts[n,3], scale_fac) flux loops given in memory.'''
        Ms_loc = c_void_p()
        Msc_loc = c_void_p()
        if sensors is not None:
            fil_ptr = numpy.zeros(len(sensors) + 1, dtype=numpy.int32)
            for k, (p, _) in enumerate(sensors):
                fil_ptr[k + 1] = fil_ptr[k] + len(p)
            P = numpy.ascontiguousarray(numpy.vstack([numpy.asarray(p, dtype=numpy.float64) for p, _ in sensors]))
            sf = numpy.array([s for _, s in sensors] + [0.0], dtype=numpy.float64)
            sensor_loc = c_void_p()
            _check(b200_set_sensors(self.tw_obj, len(sensors), fil_ptr, P, sf, ctypes.byref(sensor_loc)))
            _check(b200_msensor(self.tw_obj, sensor_loc, ctypes.byref(Ms_loc), ctypes.byref(Msc_loc)))
            nsens = len(sensors)
            names = ['FLOOP_%d' % k for k in range(nsens)]
        else:
            cache_string = self._oft_env.path2c("" if cache_file is None else cache_file)
            sensor_string = self._oft_env.path2c("none" if sensor_file is None else sensor_file)
            nsensors = c_int()
            njumpers = c_int()
            sensor_loc = c_void_p()
            error_string = self._oft_env.get_c_errorbuff()
            thincurr_Msensor(self.tw_obj, sensor_string, ctypes.byref(Ms_loc), ctypes.byref(Msc_loc), ctypes.byref(nsensors),
                             ctypes.byref(njumpers), ctypes.byref(sensor_loc), cache_string, error_string)
            if error_string.value != b'':
                raise Exception(error_string.value.decode())
            nsens = nsensors.value
            names = []
            for i in range(nsens):
                sensor_name = ctypes.create_string_buffer(b"", 40)
                error_string = self._oft_env.get_c_errorbuff()
                thincurr_get_sensor_name(sensor_loc, c_int(i + 1), sensor_name, error_string)
                if error_string.value != b'':
                    raise Exception(error_string.value.decode())
                names.append(sensor_name.value.decode().strip())
        return numpy.ctypeslib.as_array(ctypes.cast(Ms_loc, c_double_ptr), shape=(self.nelems, nsens)), \
            numpy.ctypeslib.as_array(ctypes.cast(Msc_loc, c_double_ptr), shape=(self.n_icoils, nsens)), \
            {'names': names, 'ptr': sensor_loc}

    def compute_Rmat(self, copy_out=None):
        '''! Resistance matrix as `scipy.sparse.csr_array` in `self.Rmat` (reference: _core.py:491-513)'''
        kr_loc = c_int_ptr()
        lc_loc = c_int_ptr()
        mat_loc = c_double_ptr()
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_Rmat(self.tw_obj, ctypes.byref(kr_loc), ctypes.byref(lc_loc), ctypes.byref(mat_loc), error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())
        kr = numpy.ctypeslib.as_array(kr_loc, shape=(self.nelems + 1,))
        nnz = kr[-1] - 1
        lc = numpy.ctypeslib.as_array(lc_loc, shape=(nnz,))
        data = numpy.ctypeslib.as_array(mat_loc, shape=(nnz,))
        self.Rmat = scipy.sparse.csr_array((data.copy(), lc - 1, kr - 1), shape=(self.nelems, self.nelems))

    def cross_coupling(self, model2, cache_file=None):
        '''! Mutual inductance between this and another ThinCurr model `(self.nelems, model2.nelems)`'''
        Mmat = numpy.zeros((self.nelems, model2.nelems), dtype=numpy.float64)
        cache_string = self._oft_env.path2c("" if cache_file is None else cache_file)
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_cross_coupling(self.tw_obj, model2.tw_obj, Mmat, cache_string, error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())
        return Mmat

    def get_eta_values(self):
        eta = numpy.zeros((self.nregs,), dtype=numpy.float64)
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_get_eta(self.tw_obj, eta, error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())
        return eta

    def set_eta_values(self, eta_values=None, eta_surf=None, eta_vol=None, thickness=None):
        if eta_values is not None:
            eta_surf = eta_values
        arrs = []
        for name, v in (('eta_surf', eta_surf), ('eta_vol', eta_vol), ('thickness', thickness)):
            if v is None:
                arrs.append(None)
                continue
            v = numpy.ascontiguousarray(v, dtype=numpy.float64)
            if v.shape[0] != self.nregs:
                raise IndexError('Incorrect shape of "{0}", should be [nregs]'.format(name))
            if numpy.any(v <= 0.0):
                raise ValueError('All values in "{0}" must be > 0'.format(name))
            arrs.append(v)
        ptrs = [a.ctypes.data_as(c_void_p) if a is not None else c_void_p() for a in arrs]
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_set_eta(self.tw_obj, ptrs[0], ptrs[1], ptrs[2], error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())

    # ---- sharded / device-resident builds (B200-native extension) --------------------------------
    def shard_rows(self, nshards, shard):
        '''! Reference (0-based) DOF ids of the rows owned by `shard` out of `nshards`.'''
        n = c_int()
        _check(b200_plan(self.tw_obj, nshards, shard, ctypes.byref(n)))
        rows = numpy.zeros(max(n.value, 1), dtype=numpy.int32)
        _check(b200_shard_rows(self.tw_obj, nshards, shard, rows))
        return rows[:n.value]

    def compute_Lmat_shard(self, nshards, shard, out, stream=None, stats=False):
        '''! Build the rows of L owned by `shard` into the CUDA tensor `out[nrows, ld]` (float64,
        row r = Lmat[:, row_ids[r]]).  Asynchronous on `stream` unless `stats`.'''
        st = numpy.zeros(8, dtype=numpy.int64)
        sptr = c_void_p(stream) if stream else c_void_p()
        _check(b200_Lmat_shard(self.tw_obj, nshards, shard, c_void_p(out.data_ptr()), out.stride(0), sptr,
                               st.ctypes.data_as(c_void_p) if stats else None))
        return st

    # symmetric multi-device build: upper trapezoid per shard + one exchange of the transposed blocks
    def shard_rows_sym(self, nshards, shard):
        '''! Rows (reference 0-based DOF ids) of `shard` in the symmetric partition, which equalises the work of
        the upper trapezoid (rows of the shard x DOFs of this and later shards).'''
        n = c_int()
        _check(b200_shard_rows_sym(self.tw_obj, nshards, shard, ctypes.byref(n), None))
        rows = numpy.zeros(max(n.value, 1), dtype=numpy.int32)
        _check(b200_shard_rows_sym(self.tw_obj, nshards, shard, None, rows.ctypes.data_as(c_void_p)))
        return rows[:n.value]

    def compute_Lmat_shard_sym(self, nshards, shard, out, stream=None, stats=False):
        '''! Build rows `shard_rows_sym(nshards, shard)` of L into the CUDA tensor `out[nrows, ld]` for the columns of
        this and later shards; the columns of earlier shards' DOFs stay zero until `exchange_symmetric`.'''
        st = numpy.zeros(8, dtype=numpy.int64)
        sptr = c_void_p(stream) if stream else c_void_p()
        _check(b200_Lmat_shard_sym(self.tw_obj, nshards, shard, c_void_p(out.data_ptr()), out.stride(0), sptr,
                                   st.ctypes.data_as(c_void_p) if stats else None))
        return st

    def sym_computed_mask(self, nshards, shard, row_ids=None):
        '''! Boolean mask `[nrows, nelems]` of the entries `compute_Lmat_shard_sym` evaluates in place for `shard`: its
        diagonal block and, of every block shared with another shard, the checkerboard half that belongs to its rows.'''
        from .._interface import b200_dof_patches
        pat = numpy.zeros(self.nelems, dtype=numpy.int32)
        _check(b200_dof_patches(self.tw_obj, nshards, pat))
        if row_ids is None:
            row_ids = [self.shard_rows_sym(nshards, s) for s in range(nshards)]
        owner = numpy.zeros(self.nelems, dtype=numpy.int32)
        for s, ids in enumerate(row_ids):
            owner[ids] = s
        pa = pat[row_ids[shard]][:, None]
        pb = pat[None, :]
        mine = (((pa + pb) & 1) == 0) == (pa < pb)
        return numpy.where(owner[None, :] == shard, True, mine)

    def exchange_symmetric(self, out, nshards, shard, group=None, row_ids=None):
        '''! Host-side reference of the exchange after `compute_Lmat_shard_sym` with torch.distributed collectives (any
        backend; the GPU path is `exchange_symmetric_peer`, inside the library): the entries of `out[nrows, ld]` that
        another shard evaluated are the transposes of entries of that shard's rows (thin_wall.F90:1146-1151 mirrors the
        same way).  Every rank sends each peer the block of its rows against the peer's DOFs.'''
        import torch
        import torch.distributed as dist
        if nshards == 1:
            return
        if row_ids is None:
            row_ids = [self.shard_rows_sym(nshards, s) for s in range(nshards)]
        ids = [torch.as_tensor(numpy.ascontiguousarray(r, dtype=numpy.int64), device=out.device) for r in row_ids]
        nmine = len(row_ids[shard])
        mine = out[:nmine]
        mask = torch.as_tensor(self.sym_computed_mask(nshards, shard, row_ids), device=out.device)
        for d in range(1, nshards):
            dst, src = (shard + d) % nshards, (shard - d) % nshards
            send = mine.index_select(1, ids[dst]).contiguous()                      # [my rows, DOFs of dst]
            recv = torch.empty((len(row_ids[src]), nmine), dtype=out.dtype, device=out.device)  # [rows of src, my DOFs]
            ops = [dist.P2POp(dist.isend, send, dst, group=group), dist.P2POp(dist.irecv, recv, src, group=group)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            cur = mine.index_select(1, ids[src])
            keep = mask.index_select(1, ids[src])
            mine.index_copy_(1, ids[src], torch.where(keep, cur, recv.t()))

    def compute_Lmat_shard_host(self, nshards, shard, out, stats=False):
        '''! End-to-end variant: host mesh -> device build -> host rows (`out` is a numpy array).'''
        st = numpy.zeros(8, dtype=numpy.int64)
        _check(b200_Lmat_shard_host(self.tw_obj, nshards, shard, out.ctypes.data_as(c_void_p), out.shape[1],
                                    st.ctypes.data_as(c_void_p) if stats else None))
        return st

    def compute_Lmat_block(self, row_ids, col_ids, out, stream=None):
        '''! Dense block `out[i, j] = L[row_ids[i], col_ids[j]]` (vertex/hole DOFs, 0-based reference ids) into the CUDA
        tensor `out[len(row_ids), ld >= len(col_ids)]` (float64).  Asynchronous on `stream`.'''
        rows = numpy.ascontiguousarray(row_ids, dtype=numpy.int32)
        cols = numpy.ascontiguousarray(col_ids, dtype=numpy.int32)
        sptr = c_void_p(stream) if stream else c_void_p()
        _check(b200_Lmat_block(self.tw_obj, len(rows), rows, len(cols), cols, c_void_p(out.data_ptr()), out.stride(0), sptr))

    # ---- HODLR dense-block builders, matrix-free apply, reduced model ---------------------------------
    @staticmethod
    def _out_ptr(out):
        """host numpy array or CUDA tensor -> (pointer, leading dimension)"""
        if isinstance(out, numpy.ndarray):
            return out.ctypes.data_as(c_void_p), out.strides[-2] // 8
        return c_void_p(out.data_ptr()), out.stride(-2)

    def compute_Lmatblock(self, row_pts, col_pts, out=None, col_model=None, stream=None):
        '''! (extension) `tw_compute_Lmatblock` (thin_wall_hodlr.F90:289-404): dense block between two VERTEX blocks
        (0-based mesh vertex ids), `out[a, b] = Lmat(col_pts[b], row_pts[a])`, row block's cells analytic for near pairs.
        `out`: numpy array or CUDA tensor `[len(row_pts), ld >= len(col_pts)]` (allocated on the host when None).'''
        from .._interface import b200_Lmatblock
        rows = numpy.ascontiguousarray(row_pts, dtype=numpy.int32)
        cols = numpy.ascontiguousarray(col_pts, dtype=numpy.int32)
        if out is None:
            out = numpy.zeros((len(rows), len(cols)))
        ptr, ld = self._out_ptr(out)
        _check(b200_Lmatblock(self.tw_obj, col_model.tw_obj if col_model is not None else c_void_p(), len(rows), rows, len(cols), cols,
                              ptr, ld, c_void_p(stream) if stream else c_void_p()))
        return out

    def compute_LmatHole(self, out=None, stream=None):
        '''! (extension) `tw_compute_LmatHole(self, self)` (thin_wall_hodlr.F90:136-285): `out[h, :] = Lmat(:, h)` for the
        hole and V-coil columns, `[nholes + n_vcoils, nelems]`.'''
        from .._interface import b200_LmatHole
        if out is None:
            out = numpy.zeros((self.nholes + self.n_vcoils, self.nelems))
        ptr, ld = self._out_ptr(out)
        _check(b200_LmatHole(self.tw_obj, ptr, ld, c_void_p(stream) if stream else c_void_p()))
        return out

    def compute_Bops_block(self, row_pts, col_pts, direction=-1, out=None, stream=None):
        '''! (extension) `tw_compute_Bops_block` (thin_wall_hodlr.F90:580-691): `out[a, b] = Bop(col_pts[b], row_pts[a])`
        for Cartesian component `direction` (0, 1, 2), or all three (`direction < 0`, `out[3, nrows, ncols]`).'''
        from .._interface import b200_Bops_block
        rows = numpy.ascontiguousarray(row_pts, dtype=numpy.int32)
        cols = numpy.ascontiguousarray(col_pts, dtype=numpy.int32)
        if out is None:
            out = numpy.zeros((3, len(rows), len(cols)) if direction < 0 else (len(rows), len(cols)))
        ptr, ld = self._out_ptr(out)
        _check(b200_Bops_block(self.tw_obj, len(rows), rows, len(cols), cols, int(direction), ptr, ld,
                               c_void_p(stream) if stream else c_void_p()))
        return out

    def cross_eval(self, model2, field, counts=None):
        '''! Flux induced on `model2` by current fields on this model (`tw_compute_Lmat_MF`, thin_wall.F90:1190-1414;
        reference `_core.py:533-549`): `field [nrhs, nelems]` -> `[nrhs, model2.nelems]`.'''
        field = numpy.atleast_2d(field)
        nrhs = field.shape[0]
        if field.shape[1] != self.nelems:
            raise IndexError('Incorrect shape of "field", should be [:,nelems]')
        vec_out = numpy.zeros((nrhs, model2.nelems), dtype=numpy.float64)
        vec_in = numpy.ascontiguousarray(field.copy(), dtype=numpy.float64)
        if counts is not None:
            from .._interface import b200_cross_eval
            _check(b200_cross_eval(self.tw_obj, model2.tw_obj, c_int(nrhs), vec_in, vec_out, counts.ctypes.data_as(c_void_p)))
            return vec_out
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_cross_eval(self.tw_obj, model2.tw_obj, c_int(nrhs), vec_in, vec_out, error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())
        return vec_out

    def build_reduced_model(self, basis_set, filename='tCurr_reduced.h5', compute_B=False, sensor_obj=None):
        '''! Project the model onto a basis of currents (`tw_reduce_model`, thin_wall_solvers.F90:1180-1359; reference
        `_core.py:717-739`) and write the reduced-model file.  Returns the file name (reading it back needs h5py, as in the
        reference's `ThinCurr_reduced`).'''
        basis_set = numpy.ascontiguousarray(basis_set, dtype=numpy.float64)
        sensor_ptr = sensor_obj['ptr'] if sensor_obj is not None else c_void_p()
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_reduce_model(self.tw_obj, self._oft_env.path2c(filename), c_int(basis_set.shape[0]), basis_set, c_bool(compute_B),
                              sensor_ptr, c_void_p(), error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())
        return filename

    def compute_Bel_shard(self, nshards, shard, out, stream=None):
        '''! Rows of the B operator into the CUDA tensor `out[3, np, nrows]`.'''
        sptr = c_void_p(stream) if stream else c_void_p()
        _check(b200_Bel_shard(self.tw_obj, nshards, shard, c_void_p(out.data_ptr()), sptr))

    def stream_plan(self):
        '''! Banded plan of the streamed single-device build behind `compute_Lmat()` (thincurr_b200_stream_plan): returns
        `(band_ref_ptr, band_patch_ptr)` -- band b = reference DOF ids `[band_ref_ptr[b], band_ref_ptr[b+1])` = patches
        `[band_patch_ptr[b], band_patch_ptr[b+1])` -- or `None` when this model is built the ordinary way.'''
        from .._interface import b200_stream_plan
        nb = c_int()
        ref, pat = numpy.zeros(33, dtype=numpy.int32), numpy.zeros(33, dtype=numpy.int32)
        _check(b200_stream_plan(self.tw_obj, ctypes.byref(nb), ref.ctypes.data_as(c_void_p), pat.ctypes.data_as(c_void_p)))
        if nb.value == 0:
            return None
        return ref[:nb.value + 1].copy(), pat[:nb.value + 1].copy()

    def plan_info(self):
        '''! Patch/chunk/tile counts of the owner-computes plan (see thincurr_b200_plan_info).'''
        from .._interface import b200_plan_info
        info = numpy.zeros(8, dtype=numpy.int64)
        _check(b200_plan_info(self.tw_obj, info))
        from .._interface import b200_model_bytes
        d = dict(zip(('patch_size', 'npatch', 'nchunk', 'patch_cells', 'ntiles', 'chunk_pairs', 'cell_pairs', 'nvert_patch'),
                     [int(v) for v in info]))
        d['model_bytes'] = int(b200_model_bytes(self.tw_obj))
        return d

    def pair_stats(self):
        '''! iquad histogram [19] and number of ordered pairs visited by the reference loop nest.'''
        hist = numpy.zeros(19, dtype=numpy.int64)
        vis = ctypes.c_int64()
        _check(b200_pair_stats(self.tw_obj, hist, ctypes.byref(vis)))
        return hist, vis.value

    def get_model_arrays(self):
        '''! Internal model arrays for tests: pmap, lc (oriented), kfh, lfh, qbasis, ca'''
        pmap = numpy.zeros(self.np, dtype=numpy.int32)
        lc = numpy.zeros((self.nc, 3), dtype=numpy.int32)
        kfh = numpy.zeros(self.nc + 1, dtype=numpy.int32)
        qb = numpy.zeros((self.nc, 3, 3))
        ca = numpy.zeros(self.nc)
        nfh = b200_get_model(self.tw_obj, pmap.ctypes.data_as(c_void_p), lc.ctypes.data_as(c_void_p),
                             kfh.ctypes.data_as(c_void_p), None, qb.ctypes.data_as(c_void_p), ca.ctypes.data_as(c_void_p))
        lfh = numpy.zeros((max(nfh, 1), 2), dtype=numpy.int32)
        b200_get_model(self.tw_obj, None, None, None, lfh.ctypes.data_as(c_void_p), None, None)
        return dict(pmap=pmap, lc=lc, kfh=kfh, lfh=lfh[:nfh], qbasis=qb, ca=ca)

    def model_hashes(self):
        a, b = ctypes.c_int32(), ctypes.c_int32()
        b200_hashes(self.tw_obj, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    # ---- dense apply / leading eigenmodes on the device-resident matrix ---------------------------
    def apply_Lmat(self, vals):
        '''! y = L x with the dense matrix of `compute_Lmat` (reference: thincurr_apply_Lmat, thincurr_f.F90:470-497);
        returns a new array.'''
        out = numpy.ascontiguousarray(vals, dtype=numpy.float64).copy()
        if out.shape[0] != self.nelems:
            raise IndexError('Incorrect shape of "vals", should be [nelems]')
        thincurr_apply_Lmat(self.tw_obj, out, self.Lmat_hodlr)
        return out

    def get_eigs(self, neigs, direct=False):
        '''! Leading L/R eigenmodes (reference: _core.py:563-580).  Iterative path only (Lanczos on the device-resident
        matrix); `direct=True` is refused by the library.

        @result Eigenvalues `(neigs)`, eigenvectors `(neigs,:)`'''
        eig_vals = numpy.zeros((neigs,), dtype=numpy.float64)
        eig_vecs = numpy.zeros((neigs, self.nelems), dtype=numpy.float64)
        error_string = self._oft_env.get_c_errorbuff()
        thincurr_eigenvalues(self.tw_obj, c_bool(direct), c_int(neigs), eig_vals, eig_vecs, c_void_p(), error_string)
        if error_string.value != b'':
            raise Exception(error_string.value.decode())
        return eig_vals, eig_vecs

    def get_eigs_sharded(self, neigs, apply, tol=1e-10, max_dim=400):
        '''! (extension) Leading L/R eigenmodes with the mat-vec supplied by the caller: `apply(x) -> y` on host vectors
        of `nelems` doubles, e.g. `rows_apply` on every rank's resident row block + an all-gather of y.  Needs
        `compute_Rmat`.  Returns eigenvalues, eigenvectors `(neigs,:)` and the number of mat-vecs.'''
        n = self.nelems

        def _cb(_user, xp, yp):
            try:
                x = numpy.ctypeslib.as_array(xp, shape=(n,))
                y = numpy.ctypeslib.as_array(yp, shape=(n,))
                y[:] = apply(x)
                return 0
            except Exception:  # pragma: no cover
                import traceback
                traceback.print_exc()
                return 1
        cb = B200_APPLY_FN(_cb)
        eig_vals = numpy.zeros((neigs,), dtype=numpy.float64)
        eig_vecs = numpy.zeros((neigs, n), dtype=numpy.float64)
        napp = c_int()
        _check(b200_lr_eigs(self.tw_obj, neigs, tol, max_dim, cb, None, eig_vals, eig_vecs, ctypes.byref(napp)))
        return eig_vals, eig_vecs, napp.value

    @staticmethod
    def rows_apply(rows_ptr, ld, nrows, n, x_ptr, y_ptr, stream=None):
        '''! (extension) y[nrows] = rows[nrows][ld] . x[n] on the current device (device pointers as ints).'''
        _check(b200_rows_apply(c_void_p(rows_ptr), ld, nrows, n, c_void_p(x_ptr), c_void_p(y_ptr), c_void_p(stream) if stream else c_void_p()))

    # ---- multi-device data plane in the library (include/thincurr_b200.h block 3) ---------------------
    @staticmethod
    def device_alloc(nbytes):
        '''! (extension) cudaMalloc'ed block on the current device that other ranks can map (returns the pointer as int).'''
        p = c_void_p()
        _check(b200_device_alloc(int(nbytes), ctypes.byref(p)))
        return p.value

    @staticmethod
    def device_free(ptr):
        _check(b200_device_free(c_void_p(ptr)))

    @staticmethod
    def ipc_export(ptr):
        h = ctypes.create_string_buffer(64)
        _check(b200_ipc_export(c_void_p(ptr), h))
        return h.raw

    @staticmethod
    def ipc_open(handle):
        p = c_void_p()
        _check(b200_ipc_open(ctypes.create_string_buffer(handle, 64), ctypes.byref(p)))
        return p.value

    @staticmethod
    def ipc_close(ptr):
        _check(b200_ipc_close(c_void_p(ptr)))

    def exchange_symmetric_peer(self, out_ptr, ld, nshards, shard, peer_ptrs, stream=None):
        '''! (extension) Complete the rows built by `compute_Lmat_shard_sym` by READING the transposed blocks from the
        earlier shards' row blocks (`peer_ptrs[s]`: device pointer readable from this device -- a peer device of this
        process or a cudaIpc mapping -- for every s < shard) over NVLink, inside the library
        (thincurr_b200_Lmat_exchange).  Asynchronous on `stream`; the caller orders it after the peers' builds.'''
        arr = (c_void_p * nshards)(*[c_void_p(p) if p else c_void_p() for p in peer_ptrs])
        _check(b200_Lmat_exchange(self.tw_obj, nshards, shard, c_void_p(out_ptr), ld, arr, c_void_p(stream) if stream else c_void_p()))

    def gather_full(self, nshards, sym, shard_ptrs, ld_src, full_ptr, ld_full, stream=None):
        '''! (extension) One gather over NVLink: the rows of all shards (device pointers readable from this device) into
        the full matrix `full[nelems][ld_full]` on the current device, reference row order.'''
        arr = (c_void_p * nshards)(*[c_void_p(p) if p else c_void_p() for p in shard_ptrs])
        _check(b200_Lmat_gather(self.tw_obj, nshards, 1 if sym else 0, arr, ld_src, c_void_p(full_ptr), ld_full,
                                c_void_p(stream) if stream else c_void_p()))

    def rows_to_host(self, nshards, shard, sym, rows_ptr, ld, full_host):
        '''! (extension) Stream this shard's rows (device) into the host matrix `full_host[nelems, ld_full]` (numpy).'''
        if full_host is None:   # stream only (rows pass through the pinned staging buffers and are dropped)
            _check(b200_rows_to_host(self.tw_obj, nshards, shard, 1 if sym else 0, c_void_p(rows_ptr), ld, c_void_p(), ld))
            return
        _check(b200_rows_to_host(self.tw_obj, nshards, shard, 1 if sym else 0, c_void_p(rows_ptr), ld,
                                 full_host.ctypes.data_as(c_void_p), full_host.shape[1]))

    def save_Lmat_begin(self, path):
        '''! (extension) Header + size of an `Lmat.save` cache file that shards then fill concurrently.'''
        _check(b200_Lmat_save_begin(self.tw_obj, self._oft_env.path2c(path)))

    def save_Lmat_rows(self, path, nshards, shard, sym, rows_ptr, ld):
        _check(b200_Lmat_save_rows(self.tw_obj, self._oft_env.path2c(path), nshards, shard, 1 if sym else 0, c_void_p(rows_ptr), ld))

    def release_device(self):
        '''! (extension) Free the device-side state of the model (plan mirrors, row scratch).'''
        _check(b200_release_device(self.tw_obj))

    # ---- out of scope (downstream consumers of the operators) ----------------------
    def _oos(self, *a, **k):
        raise NotImplementedError('Not part of the B200 operator-build backend; use the reference library for this step')

    compute_freq_response = run_td = plot_td = _oos
    setup_io = save_current = save_scalar = reconstruct_current = reconstruct_Bfield = get_regmat = _oos
