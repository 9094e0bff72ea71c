#!/usr/bin/env python
"""Dev-time tool: convert the reference's ThinCurr fixture meshes (native HDF5) into small
.npz fixtures under tests/golden/ (no HDF5 library exists on the GPU box and
/root/reference is absent there).  Pure format conversion -- values are unchanged.

Sources (read-only): /root/reference/src/tests/physics/tw_test-{plate,cyl,torus,passive}.h5
                     /root/reference/src/examples/ThinCurr/{cyl,torus,ports}/thincurr_ex-*.h5
The goldens themselves (eigenvalues / FR signals) are constants quoted from
src/tests/physics/test_ThinCurr.py and live in tests/golden/goldens.json.
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from h5min import load_native_mesh

REF = '/root/reference/src'
MESHES = {
    'plate': REF + '/tests/physics/tw_test-plate.h5',
    'cyl': REF + '/tests/physics/tw_test-cyl.h5',
    'torus': REF + '/tests/physics/tw_test-torus.h5',
    'passive': REF + '/tests/physics/tw_test-passive.h5',
    'ex_cyl': REF + '/examples/ThinCurr/cyl/thincurr_ex-cyl.h5',
    'ex_torus': REF + '/examples/ThinCurr/torus/thincurr_ex-torus.h5',
    'ex_ports': REF + '/examples/ThinCurr/ports/thincurr_ex-ports.h5',
}
GOLDENS = {  # src/tests/physics/test_ThinCurr.py
    'eig_plate': {'line': 984, 'vals': [9.735667E-3, 6.532314E-3, 6.532201E-3, 5.251598E-3], 'tol': 1e-5},
    'eig_cyl': {'line': 1029, 'vals': [2.657195E-2, 1.248071E-2, 1.247103E-2, 1.200566E-2], 'tol': 1e-5, 'jumper_start': 2},
    'eig_torus': {'line': 1076, 'vals': [4.751344E-2, 2.564491E-2, 2.555695E-2, 2.285850E-2], 'tol': 1e-5},
    'eig_passive': {'line': 1164, 'vals': [1.503561E-1, 6.420533E-2, 3.188782E-2, 2.941118E-2], 'tol': 1e-5,
                    'vcoils': [[0.5, 0.1], [0.5, 0.05], [0.5, -0.05], [0.5, -0.1]], 'eta': 1e4},
    'fr_plate': {'line': 1004, 'real': [6.807649E-2, 7.207748E-2], 'imag': [-3.011666E-3, -2.177010E-3], 'tol': 1e-4,
                 'icoils': [[0.5, 0.1]], 'floops': [[0.5, -0.05], [0.5, -0.1]], 'freq': 5e3},
    'fr_cyl': {'line': 1050, 'real': [6.118337E-2, 4.356188E-3], 'imag': [-1.911861E-3, -2.283493E-3], 'tol': 1e-4,
               'icoils': [[1.1, 0.25], [1.1, -0.25]], 'floops': [[0.9, 0.5], [0.9, 0.0]], 'freq': 5e3, 'jumper_start': 2},
    'fr_torus': {'line': 1099, 'real': [-2.807955E-3, -1.196091E-4], 'imag': [-1.869732E-3, -1.248642E-4], 'tol': 1e-4,
                 'icoils': [[1.5, 0.5], [1.5, -0.5]], 'floops': [[1.4, 0.0], [0.6, 0.0]], 'freq': 5e3},
    'fr_passive': {'line': 1189, 'real': [1.947713E-1, 1.990873E-1], 'imag': [-2.175942E-4, -1.560726E-4], 'tol': 1e-4,
                   'icoils': [[0.5, 0.1]], 'vcoils': [[0.5, 0.0]], 'floops': [[0.5, -0.05], [0.5, -0.1]], 'freq': 5e3, 'eta': 1e4},
}

if __name__ == '__main__':
    out = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out, exist_ok=True)
    for name, fn in MESHES.items():
        m = load_native_mesh(fn)
        d = dict(r=m['r'], lc=m['lc'], reg=m['reg'], n_nodesets=len(m['nodesets']), n_sidesets=len(m['sidesets']))
        for k, ns in enumerate(m['nodesets']):
            d['nodeset%d' % k] = ns
        for k, ss in enumerate(m['sidesets']):
            d['sideset%d' % k] = ss
        if m['pmap'] is not None:
            d['pmap'] = m['pmap']
        np.savez_compressed(os.path.join(out, 'mesh_%s.npz' % name), **d)
        print(name, m['r'].shape, m['lc'].shape, [len(n) for n in m['nodesets']], [len(s) for s in m['sidesets']])
    # quadrature KAT (src/tests/grid/quad_2d.tests): order, a, b?, ... kept verbatim as numbers
    rows = [l.split() for l in open(REF + '/tests/grid/quad_2d.tests').read().strip().split('\n')[1:]]
    kat = np.array([[float(x) for x in r] for r in rows])
    np.save(os.path.join(out, 'quad_2d_kat.npy'), kat)
    json.dump(GOLDENS, open(os.path.join(out, 'goldens.json'), 'w'), indent=1)


def reference_interface_symbols():
    """Names the reference's Python layer binds from liboftpy.so for ThinCurr (ThinCurr/_interface.py:17-119):
    a replacement library must export every one of them or `import OpenFUSIONToolkit.ThinCurr` raises AttributeError."""
    import re
    src = open('/root/reference/src/python/OpenFUSIONToolkit/ThinCurr/_interface.py').read()
    return sorted(set(re.findall(r'oftpy_lib\.(\w+)', src)))


if __name__ == '__main__' and '--symbols' in sys.argv:
    names = reference_interface_symbols()
    json.dump({'source': 'src/python/OpenFUSIONToolkit/ThinCurr/_interface.py:17-119', 'names': names},
              open(os.path.join(ROOT, 'tests', 'golden', 'ref_interface_symbols.json'), 'w'), indent=1)
    print('wrote %d names' % len(names))
