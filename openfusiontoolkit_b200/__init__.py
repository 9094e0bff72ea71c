"""thincurr-b200: B200-native dense operator builds for ThinCurr (Open FUSION Toolkit).

Only the hot path of the reference is provided: the O(N^2) thin-wall inductance operator
build and the coil / sensor / B-field operators, behind the reference's ThinCurr Python API
(`ThinCurr.setup_model / compute_Lmat / compute_Bmat / compute_Mcoil / compute_Msensor /
compute_Rmat / cross_coupling`).  The compute path is hand-written CUDA for sm_100a reached
through the C ABI in include/thincurr_b200.h; there is no CPU fallback.
"""
__all__ = ['OFT_env']


def __getattr__(name):
    # lazy so that `python -m openfusiontoolkit_b200.build` can run before the library exists
    if name == 'OFT_env':
        from ._core import OFT_env
        return OFT_env
    raise AttributeError(name)
