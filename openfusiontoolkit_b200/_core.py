"""Runtime environment object (mirror of OpenFUSIONToolkit._core.OFT_env, _core.py:24-140,
reduced to what the ThinCurr operator-build path uses)."""
import ctypes
import numpy
from ._interface import oft_init, oftpy_set_debug, oftpy_set_nthreads


class OFT_env():
    '''! Execution environment.  `nthreads` is accepted for API compatibility; the operator
    builds run on the CUDA devices visible to the process (THINCURR_B200_NDEV caps how many).'''
    _initialized = False

    def __init__(self, debug_level=0, nthreads=-1, unique_tempfiles='global', abort_callback=True, quiet=True):
        self.nthreads = nthreads
        self.debug_level = debug_level
        self.oft_in_groups = {}
        slens = numpy.zeros((4,), dtype=numpy.int32)
        oft_init(int(nthreads), bool(quiet), b'', slens, None)
        self.oft_mpi_plen, self.oft_slen, self.oft_path_slen, self.oft_error_slen = [int(v) for v in slens]
        oftpy_set_debug(int(debug_level))
        OFT_env._initialized = True

    def update_oft_in(self):
        pass

    def set_debug_level(self, debug_level):
        oftpy_set_debug(int(debug_level))

    def set_num_threads(self, nthreads):
        oftpy_set_nthreads(int(nthreads))

    def path2c(self, path):
        if len(path) >= self.oft_path_slen:
            raise ValueError('Path "{0}" exceeds the maximum path length'.format(path))
        return ctypes.c_char_p(path.encode())

    def get_c_errorbuff(self):
        return ctypes.create_string_buffer(b"", self.oft_error_slen)
