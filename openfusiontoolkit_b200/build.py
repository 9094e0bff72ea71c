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


TEST_LIB = os.path.join(HERE, 'libthincurr_b200_test.so')


def build_test(force=False):
    """Test build: the same sources with -DTW_TEST_HOOKS (kernel probes of csrc/tw_probe.cu, the switches that force the
    rare contraction paths).  Loaded only by tests/ through openfusiontoolkit_b200._testlib; never by the product path."""
    if not force and os.path.exists(TEST_LIB) and os.path.getmtime(TEST_LIB) >= os.path.getmtime(build()):
        stale = any(os.path.getmtime(os.path.join(CSRC, f)) > os.path.getmtime(TEST_LIB) for f in os.listdir(CSRC))
        if not stale:
            return TEST_LIB
    return build(defs=('TW_TEST_HOOKS',), out=TEST_LIB)


if __name__ == '__main__':
    defs = [a[2:] for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith('-o')]
    if '--test' in sys.argv:
        print('built', build_test(force='--force' in sys.argv))
        sys.exit(0)
    print('built', build(force='--force' in sys.argv, verbose='-v' in sys.argv, defs=defs, out=outs[0] if outs else None))
