#!/usr/bin/env python
"""Dev-time tool: place the UNMODIFIED pure-Python layer of the reference (src/python/OpenFUSIONToolkit) under
baseline/_ref/ (git-ignored, travels to the GPU box with gpurun) with `liboftpy.so` / `liboft_triangle.so` pointing at
libthincurr_b200.so -- the loader switch of INTEGRATION.md section B.  The reference itself cannot be installed in this
image (`pip install /root/reference` needs cmake + a Fortran compiler + HDF5; none exist), so only its Python layer is
used, driving this repo's library.  Nothing is copied into the tracked tree."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference/src/python/OpenFUSIONToolkit'
DST = os.path.join(ROOT, 'baseline', '_ref', 'OpenFUSIONToolkit')


def install():
    if not os.path.isdir(SRC):
        print('reference not present: nothing installed')
        return False
    if os.path.isdir(DST):
        os.system('chmod -R u+w %s' % DST)
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns('CMakeLists.txt', '__pycache__'))
    os.system('chmod -R u+w %s' % DST)
    lib = os.path.join('..', '..', '..', 'openfusiontoolkit_b200', 'libthincurr_b200.so')
    for name in ('liboftpy.so', 'liboft_triangle.so'):
        os.symlink(lib, os.path.join(DST, name))
    print('installed', DST)
    return True


if __name__ == '__main__':
    sys.exit(0 if install() else 1)
