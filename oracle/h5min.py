"""Minimal read-only HDF5 parser -- TEST INFRASTRUCTURE (oracle side).

Reads the subset of HDF5 that the reference's native ThinCurr mesh files use
(superblock v0, v1 object headers, symbol-table groups, contiguous uncompressed
little-endian int/float datasets; layout written by
src/python/OpenFUSIONToolkit/util.py:62-94 `write_native_mesh`).  No h5py / libhdf5
exists in this image.  Only used by tools/make_golden.py and the oracle tests.
"""
import struct
import numpy as np


class H5:
    def __init__(self, fn):
        self.b = open(fn, 'rb').read()
        if self.b[:8] != b'\x89HDF\r\n\x1a\n' or self.b[8] != 0:
            raise ValueError('not a superblock-v0 HDF5 file: %s' % fn)
        if self.b[13] != 8 or self.b[14] != 8:
            raise ValueError('only 8-byte offsets/lengths supported')
        self.root = self._ste(56)

    def _ste(self, off):
        name_off, ohdr, cache = struct.unpack_from('<QQI', self.b, off)
        return dict(name_off=name_off, ohdr=ohdr, cache=cache)

    def _msgs(self, addr):
        ver, _, nmsg, _refc, hsize = struct.unpack_from('<BBHII', self.b, addr)
        assert ver == 1
        msgs, blocks = [], [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(msgs) < nmsg:
                t, ms, _fl = struct.unpack_from('<HHB', self.b, p)
                body = self.b[p + 8:p + 8 + ms]
                if t == 0x10:
                    o, l = struct.unpack_from('<QQ', body, 0)
                    blocks.append((o, l))
                msgs.append((t, body))
                p += 8 + ms
        return msgs

    def _group(self, ohdr):
        for t, body in self._msgs(ohdr):
            if t == 0x11:
                bt, heap = struct.unpack_from('<QQ', body, 0)
                return self._walk(bt, heap)
        return None

    def _walk(self, bt, heap):
        out = {}
        assert self.b[heap:heap + 4] == b'HEAP'
        hd = struct.unpack_from('<QQQ', self.b, heap + 8)[2]
        assert self.b[bt:bt + 4] == b'TREE'
        _ntype, level, used = struct.unpack_from('<BBH', self.b, bt + 4)
        p = bt + 8 + 16
        children = []
        for _ in range(used):
            p += 8
            children.append(struct.unpack_from('<Q', self.b, p)[0])
            p += 8
        for c in children:
            if level > 0:
                out.update(self._walk(c, heap))
            else:
                assert self.b[c:c + 4] == b'SNOD'
                n = struct.unpack_from('<H', self.b, c + 6)[0]
                for i in range(n):
                    e = self._ste(c + 8 + 40 * i)
                    nm = self.b[hd + e['name_off']:].split(b'\0', 1)[0].decode()
                    out[nm] = e['ohdr']
        return out

    def tree(self, ohdr=None, prefix=''):
        if ohdr is None:
            ohdr = self.root['ohdr']
        res = {}
        g = self._group(ohdr)
        if g is None:
            return None
        for k, v in g.items():
            sub = self.tree(v, prefix + k + '/')
            if sub is None:
                res[prefix + k] = v
            else:
                res.update(sub)
        return res

    def read(self, ohdr):
        shape = dt = layout = None
        for t, body in self._msgs(ohdr):
            if t == 1:
                ver, rank = body[0], body[1]
                off = 8 if ver == 1 else 4
                shape = struct.unpack_from('<%dQ' % rank, body, off)
            elif t == 3:
                cls = body[0] & 0xf
                size = struct.unpack_from('<I', body, 4)[0]
                signed = (body[1] >> 3) & 1
                dt = {0: ('i' if signed else 'u'), 1: 'f'}[cls] + str(size)
            elif t == 8:
                ver = body[0]
                if ver == 3:
                    assert body[1] == 1, 'only contiguous layout supported'
                    layout = struct.unpack_from('<QQ', body, 2)[0]
                else:
                    assert body[2] == 1
                    layout = struct.unpack_from('<Q', body, 8)[0]
        n = int(np.prod(shape)) if shape else 1
        if n == 0:  # zero-size dataset: no storage allocated (address undefined)
            return np.zeros(shape, dtype='<' + dt)
        a = np.frombuffer(self.b, dtype='<' + dt, count=n, offset=layout)
        return a.reshape(shape) if shape else a


def load_native_mesh(fn):
    """Return dict(r[np,3] f8, lc[nc,3] i4 0-based, reg[nc] i4, nodesets[list of 0-based arrays],
    sidesets[list of 0-based arrays], pmap or None) from a native ThinCurr mesh file."""
    f = H5(fn)
    t = f.tree()
    r = np.array(f.read(t['mesh/R']), dtype=np.float64)
    if r.shape[1] == 2:
        r = np.hstack([r, np.zeros((len(r), 1))])
    lc = np.array(f.read(t['mesh/LC']), dtype=np.int32) - 1
    reg = np.array(f.read(t['mesh/REG']), dtype=np.int32) if 'mesh/REG' in t else np.ones(len(lc), np.int32)
    out = dict(r=r, lc=lc, reg=reg, nodesets=[], sidesets=[], pmap=None)
    for key, name in (('nodesets', 'NODESET'), ('sidesets', 'SIDESET')):
        k = 1
        while 'mesh/%s%04d' % (name, k) in t:
            out[key].append(np.array(f.read(t['mesh/%s%04d' % (name, k)]), dtype=np.int32).ravel() - 1)
            k += 1
    if 'thincurr/periodicity/pmap' in t:
        out['pmap'] = np.array(f.read(t['thincurr/periodicity/pmap']), dtype=np.int32).ravel()
    return out
