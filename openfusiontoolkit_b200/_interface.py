"""ctypes binding of libthincurr_b200.so (mirror of the reference's
src/python/OpenFUSIONToolkit/_interface.py:38-120 for this path).

The library is built in-tree by `openfusiontoolkit_b200.build`; importing this module fails
loudly if it cannot be loaded -- there is no alternative compute path.
"""
import ctypes
import os
from ctypes import c_bool, c_char_p, c_double, c_int, c_int64, c_void_p

import numpy
from numpy import float64, int32

root_path = os.path.realpath(os.path.dirname(__file__))
lib_path = os.environ.get('THINCURR_B200_LIB') or os.path.join(root_path, 'libthincurr_b200.so')  # override: tuning variants
if not os.path.exists(lib_path):
    try:
        from .build import build as _build
        _build()
    except Exception as exc:  # pragma: no cover
        raise FileNotFoundError('libthincurr_b200.so is missing and could not be built: %s' % exc)
try:
    oftpy_lib = ctypes.CDLL(lib_path)
except OSError as load_error:  # pragma: no cover
    raise FileNotFoundError('Unable to load thincurr-b200 shared library') from load_error

c_int_ptr = ctypes.POINTER(c_int)
c_int_ptr_ptr = ctypes.POINTER(c_int_ptr)
c_double_ptr = ctypes.POINTER(c_double)
c_double_ptr_ptr = ctypes.POINTER(c_double_ptr)
c_void_ptr_ptr = ctypes.POINTER(c_void_p)
mu0 = numpy.pi * 4.E-7


def ctypes_numpy_array(type, ndim):
    return numpy.ctypeslib.ndpointer(dtype=type, ndim=ndim, flags='C_CONTIGUOUS')


def ctypes_subroutine(function, argtypes=None, restype=None):
    function.restype = restype
    if argtypes is not None:
        function.argtypes = argtypes
    return function


# ---- reference-compatible entry points (same argument lists as ThinCurr/_interface.py:17-120)
oft_init = ctypes_subroutine(oftpy_lib.oftpy_init, [c_int, c_bool, c_char_p, ctypes_numpy_array(int32, 1), c_void_p])
oftpy_load_xml = ctypes_subroutine(oftpy_lib.oftpy_load_xml, [c_char_p, c_void_ptr_ptr])
oftpy_set_debug = ctypes_subroutine(oftpy_lib.oftpy_set_debug, [c_int])
oftpy_set_nthreads = ctypes_subroutine(oftpy_lib.oftpy_set_nthreads, [c_int])
thincurr_setup = ctypes_subroutine(oftpy_lib.thincurr_setup,
    [c_char_p, c_int, ctypes_numpy_array(float64, 2), c_int, ctypes_numpy_array(int32, 2), ctypes_numpy_array(int32, 1),
     ctypes_numpy_array(int32, 1), c_int, c_void_ptr_ptr, ctypes_numpy_array(int32, 1), c_char_p, c_void_p])
thincurr_Lmat = ctypes_subroutine(oftpy_lib.thincurr_Lmat, [c_void_p, c_bool, c_void_ptr_ptr, c_char_p, c_char_p])
thincurr_Bmat = ctypes_subroutine(oftpy_lib.thincurr_Bmat, [c_void_p, c_void_p, c_void_ptr_ptr, c_void_ptr_ptr, c_char_p, c_char_p])
thincurr_Mcoil = ctypes_subroutine(oftpy_lib.thincurr_Mcoil, [c_void_p, c_void_ptr_ptr, c_char_p, c_char_p])
thincurr_Msensor = ctypes_subroutine(oftpy_lib.thincurr_Msensor,
    [c_void_p, c_char_p, c_void_ptr_ptr, c_void_ptr_ptr, c_int_ptr, c_int_ptr, c_void_ptr_ptr, c_char_p, c_char_p])
thincurr_get_sensor_name = ctypes_subroutine(oftpy_lib.thincurr_get_sensor_name, [c_void_p, c_int, c_char_p, c_char_p])
thincurr_cross_coupling = ctypes_subroutine(oftpy_lib.thincurr_cross_coupling,
    [c_void_p, c_void_p, ctypes_numpy_array(float64, 2), c_char_p, c_char_p])
thincurr_Rmat = ctypes_subroutine(oftpy_lib.thincurr_Rmat, [c_void_p, c_int_ptr_ptr, c_int_ptr_ptr, c_double_ptr_ptr, c_char_p])
thincurr_get_eta = ctypes_subroutine(oftpy_lib.thincurr_get_eta, [c_void_p, ctypes_numpy_array(float64, 1), c_char_p])
thincurr_set_eta = ctypes_subroutine(oftpy_lib.thincurr_set_eta, [c_void_p, c_void_p, c_void_p, c_void_p, c_char_p])

# ---- B200-native flat interface (include/thincurr_b200.h, block 2)
b200_last_error = ctypes_subroutine(oftpy_lib.thincurr_b200_last_error, [], c_char_p)
b200_device_count = ctypes_subroutine(oftpy_lib.thincurr_b200_device_count, [], c_int)
b200_destroy = ctypes_subroutine(oftpy_lib.thincurr_b200_destroy, [c_void_p])
b200_setup = ctypes_subroutine(oftpy_lib.thincurr_b200_setup,
    [c_int, ctypes_numpy_array(float64, 2), c_int, ctypes_numpy_array(int32, 2), c_void_p, c_void_p, c_int,
     ctypes_numpy_array(int32, 1), ctypes_numpy_array(int32, 1), c_int, ctypes_numpy_array(int32, 1), c_void_p,
     c_void_ptr_ptr, ctypes_numpy_array(int32, 1)], c_int)
b200_set_coils = ctypes_subroutine(oftpy_lib.thincurr_b200_set_coils,
    [c_void_p, c_int, c_int, ctypes_numpy_array(int32, 1), ctypes_numpy_array(int32, 1), ctypes_numpy_array(float64, 2),
     ctypes_numpy_array(float64, 1), ctypes_numpy_array(float64, 1), ctypes_numpy_array(float64, 1),
     ctypes_numpy_array(int32, 1), ctypes_numpy_array(int32, 1)], c_int)
b200_set_sensors = ctypes_subroutine(oftpy_lib.thincurr_b200_set_sensors,
    [c_void_p, c_int, ctypes_numpy_array(int32, 1), ctypes_numpy_array(float64, 2), ctypes_numpy_array(float64, 1),
     c_void_ptr_ptr], c_int)
b200_msensor = ctypes_subroutine(oftpy_lib.thincurr_b200_msensor, [c_void_p, c_void_p, c_void_ptr_ptr, c_void_ptr_ptr], c_int)
b200_plan = ctypes_subroutine(oftpy_lib.thincurr_b200_plan, [c_void_p, c_int, c_int, c_int_ptr], c_int)
b200_stream_plan = ctypes_subroutine(oftpy_lib.thincurr_b200_stream_plan, [c_void_p, c_int_ptr, c_void_p, c_void_p], c_int)
b200_shard_rows = ctypes_subroutine(oftpy_lib.thincurr_b200_shard_rows, [c_void_p, c_int, c_int, ctypes_numpy_array(int32, 1)], c_int)
b200_Lmat_shard = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmat_shard,
    [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p], c_int)
b200_shard_rows_sym = ctypes_subroutine(oftpy_lib.thincurr_b200_shard_rows_sym, [c_void_p, c_int, c_int, c_int_ptr, c_void_p], c_int)
b200_dof_patches = ctypes_subroutine(oftpy_lib.thincurr_b200_dof_patches, [c_void_p, c_int, ctypes_numpy_array(int32, 1)], c_int)
b200_Lmat_shard_sym = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmat_shard_sym,
    [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p], c_int)
b200_Lmat_shard_host = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmat_shard_host,
    [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p], c_int)
b200_Lmat_block = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmat_block,
    [c_void_p, c_int, ctypes_numpy_array(int32, 1), c_int, ctypes_numpy_array(int32, 1), c_void_p, c_int64, c_void_p], c_int)
b200_Lmatblock = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmatblock,
    [c_void_p, c_void_p, c_int, ctypes_numpy_array(int32, 1), c_int, ctypes_numpy_array(int32, 1), c_void_p, c_int64, c_void_p], c_int)
b200_LmatHole = ctypes_subroutine(oftpy_lib.thincurr_b200_LmatHole, [c_void_p, c_void_p, c_int64, c_void_p], c_int)
b200_Bops_block = ctypes_subroutine(oftpy_lib.thincurr_b200_Bops_block,
    [c_void_p, c_int, ctypes_numpy_array(int32, 1), c_int, ctypes_numpy_array(int32, 1), c_int, c_void_p, c_int64, c_void_p], c_int)
b200_cross_eval = ctypes_subroutine(oftpy_lib.thincurr_b200_cross_eval,
    [c_void_p, c_void_p, c_int, ctypes_numpy_array(float64, 2), ctypes_numpy_array(float64, 2), c_void_p], c_int)
b200_h5_write = ctypes_subroutine(oftpy_lib.thincurr_b200_h5_write,
    [c_char_p, c_int, ctypes.POINTER(c_char_p), ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int64), ctypes.POINTER(c_void_p)], c_int)
b200_Bel_shard = ctypes_subroutine(oftpy_lib.thincurr_b200_Bel_shard, [c_void_p, c_int, c_int, c_void_p, c_void_p], c_int)
b200_pair_stats = ctypes_subroutine(oftpy_lib.thincurr_b200_pair_stats,
    [c_void_p, numpy.ctypeslib.ndpointer(dtype=numpy.int64, ndim=1, flags='C_CONTIGUOUS'), ctypes.POINTER(c_int64)], c_int)
b200_dfma_peak = ctypes_subroutine(oftpy_lib.thincurr_b200_dfma_peak, [c_int, c_double_ptr], c_double)
b200_get_model = ctypes_subroutine(oftpy_lib.thincurr_b200_get_model,
    [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int)
b200_hashes = ctypes_subroutine(oftpy_lib.thincurr_b200_hashes, [c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)], c_int)
_f64 = numpy.ctypeslib.ndpointer(dtype=numpy.float64, flags='C_CONTIGUOUS')
_i32 = numpy.ctypeslib.ndpointer(dtype=numpy.int32, flags='C_CONTIGUOUS')
b200_launch_count = ctypes_subroutine(oftpy_lib.thincurr_b200_launch_count, [], ctypes.c_longlong)
b200_plan_info = ctypes_subroutine(oftpy_lib.thincurr_b200_plan_info, [c_void_p, numpy.ctypeslib.ndpointer(dtype=numpy.int64, flags='C_CONTIGUOUS')], c_int)
b200_model_from_tw = ctypes_subroutine(oftpy_lib.thincurr_b200_model_from_tw,
    [c_int, _f64, c_int, _i32, c_void_p, _i32, c_int, c_int, _i32, c_void_p, c_void_p, c_void_p, c_void_ptr_ptr], c_int)
b200_Lmat_host = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmat_host, [c_void_p, _f64], c_int)
b200_model_bytes = ctypes_subroutine(oftpy_lib.thincurr_b200_model_bytes, [c_void_p], c_int64)
b200_release_device = ctypes_subroutine(oftpy_lib.thincurr_b200_release_device, [c_void_p], c_int)

# ---- the remaining reference names (ThinCurr/_interface.py:21-119): native where cheap, "not provided" otherwise
thincurr_setup_io = ctypes_subroutine(oftpy_lib.thincurr_setup_io, [c_void_p, c_char_p, c_bool, c_bool, c_char_p])
thincurr_recon_curr = ctypes_subroutine(oftpy_lib.thincurr_recon_curr, [c_void_p, ctypes_numpy_array(float64, 1), ctypes_numpy_array(float64, 2), c_int])
thincurr_recon_field = ctypes_subroutine(oftpy_lib.thincurr_recon_field,
    [c_void_p, ctypes_numpy_array(float64, 1), ctypes_numpy_array(float64, 1), ctypes_numpy_array(float64, 2), c_void_p])
thincurr_save_field = ctypes_subroutine(oftpy_lib.thincurr_save_field, [c_void_p, ctypes_numpy_array(float64, 1), c_char_p])
thincurr_save_scalar = ctypes_subroutine(oftpy_lib.thincurr_save_scalar, [c_void_p, ctypes_numpy_array(float64, 1), c_char_p])
thincurr_scale_va = ctypes_subroutine(oftpy_lib.thincurr_scale_va, [c_void_p, ctypes_numpy_array(float64, 1), c_bool])
thincurr_apply_Lmat = ctypes_subroutine(oftpy_lib.thincurr_apply_Lmat, [c_void_p, ctypes_numpy_array(float64, 1), c_void_p])
thincurr_cross_eval = ctypes_subroutine(oftpy_lib.thincurr_cross_eval,
    [c_void_p, c_void_p, c_int, ctypes_numpy_array(float64, 2), ctypes_numpy_array(float64, 2), c_char_p])
thincurr_get_eta_vol = ctypes_subroutine(oftpy_lib.thincurr_get_eta_vol, [c_void_p, ctypes_numpy_array(float64, 1), c_char_p])
thincurr_get_thickness = ctypes_subroutine(oftpy_lib.thincurr_get_thickness, [c_void_p, ctypes_numpy_array(float64, 1), c_char_p])
thincurr_curr_regmat = ctypes_subroutine(oftpy_lib.thincurr_curr_regmat, [c_void_p, ctypes_numpy_array(float64, 2), c_char_p])
thincurr_eigenvalues = ctypes_subroutine(oftpy_lib.thincurr_eigenvalues,
    [c_void_p, c_bool, c_int, ctypes_numpy_array(float64, 1), ctypes_numpy_array(float64, 2), c_void_p, c_char_p])
thincurr_freq_response = ctypes_subroutine(oftpy_lib.thincurr_freq_response,
    [c_void_p, c_bool, c_int, c_double, ctypes_numpy_array(float64, 2), c_void_p, c_char_p])
thincurr_time_domain = ctypes_subroutine(oftpy_lib.thincurr_time_domain,
    [c_void_p, c_bool, c_double, c_int, c_double, c_double, c_bool, c_int, c_int, ctypes_numpy_array(float64, 1), c_void_p, c_int,
     ctypes_numpy_array(float64, 2), c_int, ctypes_numpy_array(float64, 2), c_bool, c_void_p, c_void_p, c_char_p])
thincurr_time_domain_plot = ctypes_subroutine(oftpy_lib.thincurr_time_domain_plot,
    [c_void_p, c_bool, c_bool, c_int, c_int, c_void_p, ctypes_numpy_array(float64, 2), c_int, c_void_p, c_char_p])
thincurr_reduce_model = ctypes_subroutine(oftpy_lib.thincurr_reduce_model,
    [c_void_p, c_char_p, c_int, ctypes_numpy_array(float64, 2), c_bool, c_void_p, c_void_p, c_char_p])

# ---- multi-device data plane (include/thincurr_b200.h, block 3)
_pp = ctypes.POINTER(c_void_p)
b200_device_alloc = ctypes_subroutine(oftpy_lib.thincurr_b200_device_alloc, [c_int64, c_void_ptr_ptr], c_int)
b200_device_free = ctypes_subroutine(oftpy_lib.thincurr_b200_device_free, [c_void_p], c_int)
b200_ipc_export = ctypes_subroutine(oftpy_lib.thincurr_b200_ipc_export, [c_void_p, c_char_p], c_int)
b200_ipc_open = ctypes_subroutine(oftpy_lib.thincurr_b200_ipc_open, [c_char_p, c_void_ptr_ptr], c_int)
b200_ipc_close = ctypes_subroutine(oftpy_lib.thincurr_b200_ipc_close, [c_void_p], c_int)
b200_enable_peer = ctypes_subroutine(oftpy_lib.thincurr_b200_enable_peer, [c_int], c_int)
b200_Lmat_exchange = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmat_exchange, [c_void_p, c_int, c_int, c_void_p, c_int64, _pp, c_void_p], c_int)
b200_Lmat_gather = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmat_gather, [c_void_p, c_int, c_int, _pp, c_int64, c_void_p, c_int64, c_void_p], c_int)
b200_rows_to_host = ctypes_subroutine(oftpy_lib.thincurr_b200_rows_to_host, [c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64], c_int)
b200_Lmat_save_begin = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmat_save_begin, [c_void_p, c_char_p], c_int)
b200_Lmat_save_rows = ctypes_subroutine(oftpy_lib.thincurr_b200_Lmat_save_rows, [c_void_p, c_char_p, c_int, c_int, c_int, c_void_p, c_int64], c_int)
b200_rows_apply = ctypes_subroutine(oftpy_lib.thincurr_b200_rows_apply, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p], c_int)
B200_APPLY_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_double_ptr, c_double_ptr)
b200_lr_eigs = ctypes_subroutine(oftpy_lib.thincurr_b200_lr_eigs,
    [c_void_p, c_int, c_double, c_int, B200_APPLY_FN, c_void_p, ctypes_numpy_array(float64, 1), ctypes_numpy_array(float64, 2), c_int_ptr], c_int)
