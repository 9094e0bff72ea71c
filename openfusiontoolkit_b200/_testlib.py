"""ctypes binding of the TEST build libthincurr_b200_test.so (-DTW_TEST_HOOKS): kernel-level probes
(csrc/tw_probe.cu) and the environment switches that force the rare contraction paths.  Imported only by tests/."""
import ctypes
import os
from ctypes import c_int

import numpy

from .build import TEST_LIB, build_test

if not os.path.exists(TEST_LIB):
    build_test()
test_lib = ctypes.CDLL(TEST_LIB)
_f64 = numpy.ctypeslib.ndpointer(dtype=numpy.float64, flags='C_CONTIGUOUS')
_i32 = numpy.ctypeslib.ndpointer(dtype=numpy.int32, flags='C_CONTIGUOUS')


def _sub(f, argtypes, restype=None):
    f.argtypes = argtypes
    f.restype = restype
    return f


b200_probe_pairs = _sub(test_lib.thincurr_b200_probe_pairs, [c_int, c_int, _f64, _f64, _f64, _f64, _f64, _i32], c_int)
b200_probe_phipot = _sub(test_lib.thincurr_b200_probe_phipot, [c_int, _f64, _f64, _f64], c_int)
b200_probe_rsqrt = _sub(test_lib.thincurr_b200_probe_rsqrt, [c_int, _f64, _f64], c_int)
