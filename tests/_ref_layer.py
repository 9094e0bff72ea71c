"""Helper for the drop-in tests: run a script in a fresh interpreter with the UNMODIFIED Python layer of the reference
(baseline/_ref/OpenFUSIONToolkit, staged by tools/install_reference_python.py; its loader finds `liboftpy.so` =
libthincurr_b200.so next to the package) on sys.path.  h5py does not exist in this image; the reference imports it at
module level for its plot/restart files only, so an empty stand-in module is registered first."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, 'baseline', '_ref')

PRELUDE = r'''
import sys, types, os
try:
    import h5py  # noqa: F401
except Exception:
    sys.modules['h5py'] = types.ModuleType('h5py')
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import numpy as np
from OpenFUSIONToolkit import OFT_env
from OpenFUSIONToolkit.ThinCurr import ThinCurr
import OpenFUSIONToolkit._interface as _I
assert os.path.realpath(_I.oftpy_lib._name).endswith('libthincurr_b200.so'), _I.oftpy_lib._name
GOLDEN = %r
''' % (REF_PKG, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden'))


def available():
    return os.path.exists(os.path.join(REF_PKG, 'OpenFUSIONToolkit', 'ThinCurr', '_core.py')) and \
        os.path.exists(os.path.join(REF_PKG, 'OpenFUSIONToolkit', 'liboftpy.so'))


def run(body, timeout=900, env=None):
    res = subprocess.run([sys.executable, '-c', PRELUDE + body], capture_output=True, text=True, timeout=timeout,
                         env=dict(os.environ, **(env or {})), cwd=ROOT)
    return res
