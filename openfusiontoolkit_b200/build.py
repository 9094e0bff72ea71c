"""Build libthincurr_b200.so in-tree with nvcc for sm_100a (no JIT, no torch dependency).

    python -m openfusiontoolkit_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libthincurr_b200.so')
SOURCES = ['tw_unity.cu', 'tw_setup.cpp', 'tw_io.cpp', 'tw_plan.cpp']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC,-O2,-Wall,-Wno-unused-function', '--expt-relaxed-constexpr',
              '-Xptxas', '-v', '-shared', '-cudart', 'shared']


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in os.listdir(root):
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return False


def build(force=False, verbose=False, defs=(), out=None):
    """defs/out: build a tuning variant (extra -D flags) under another file name; the product
    library is always the default build."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-D' + d for d in defs] + ['-o', out or LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, 'build.log' if out is None else os.path.basename(out) + '.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError('nvcc failed building libthincurr_b200.so')
    if verbose:
        print(log)
    return out or LIB


if __name__ == '__main__':
    defs = [a[2:] for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith('-o')]
    print('built', build(force='--force' in sys.argv, verbose='-v' in sys.argv, defs=defs, out=outs[0] if outs else None))
