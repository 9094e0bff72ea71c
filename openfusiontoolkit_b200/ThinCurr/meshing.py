"""Mesh generators: the reference's dummy mesh (ThinCurr/meshing.py:37-85) and the synthetic
tokamak-vessel meshes used by the benchmarks."""
import numpy


def build_ThinCurr_dummy(center, size=1.0, nsplit=0):
    r = numpy.array([[-size / 2.0, -size / 2.0, 0.0], [size / 2.0, -size / 2.0, 0.0], [size / 2.0, size / 2.0, 0.0],
                     [-size / 2.0, size / 2.0, 0.0], [0.0, 0.0, 0.0]]) + numpy.asarray(center, dtype=float)
    lc = numpy.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]])
    for _ in range(nsplit):
        lc_new, r_new = [], [p for p in r]
        for j in range(lc.shape[0]):
            ni = [0, 0, 0]
            cand = [(r[lc[j, 0]] + r[lc[j, 1]]) / 2.0, (r[lc[j, 1]] + r[lc[j, 2]]) / 2.0, (r[lc[j, 0]] + r[lc[j, 2]]) / 2.0]
            for k in range(3):
                for k2 in range(r.shape[0], len(r_new)):
                    if numpy.linalg.norm(r_new[k2] - cand[k]) < 1.E-10:
                        ni[k] = k2
                        break
                else:
                    r_new.append(cand[k])
                    ni[k] = len(r_new) - 1
            lc_new += [[lc[j, 0], ni[0], ni[2]], [ni[0], lc[j, 1], ni[1]], [ni[1], lc[j, 2], ni[2]], [ni[0], ni[1], ni[2]]]
        lc, r = numpy.array(lc_new), numpy.array(r_new)
    return r, lc


def build_torus_vessel(ntheta, nphi, R0=1.0, a=0.5, kappa=1.0, nports=0, jitter=0.05, seed=1234, permute_seed=None):
    """Synthetic tokamak vacuum vessel: structured (D-shaped if kappa>1) torus of ntheta x nphi
    quads split in two triangles with alternating diagonal, `nports` rectangular port cut-outs on
    the outboard side, vertex jitter +-`jitter` of the local edge length (rng(seed)).

    Returns dict(r, lc, nodesets, closures): holes = poloidal + toroidal loop (closed torus) plus
    one boundary seed vertex per port; one closure cell.  Vertex order = generator order, or a
    seeded random permutation when `permute_seed` is given.
    """
    rng = numpy.random.default_rng(seed)
    th = numpy.arange(ntheta) * 2.0 * numpy.pi / ntheta
    ph = numpy.arange(nphi) * 2.0 * numpy.pi / nphi
    TH, PH = numpy.meshgrid(th, ph, indexing='ij')
    dth, dph = 2.0 * numpy.pi / ntheta, 2.0 * numpy.pi / nphi
    TH = TH + jitter * dth * rng.uniform(-1.0, 1.0, TH.shape)
    PH = PH + jitter * dph * rng.uniform(-1.0, 1.0, PH.shape)
    Rm = R0 + a * numpy.cos(TH)
    r = numpy.stack([Rm * numpy.cos(PH), Rm * numpy.sin(PH), kappa * a * numpy.sin(TH)], -1).reshape(-1, 3)
    vid = lambda i, j: (i % ntheta) * nphi + (j % nphi)
    # port cut-outs: blocks of quads centred on the outboard midplane (theta = 0)
    removed = numpy.zeros((ntheta, nphi), dtype=bool)
    pw_t, pw_p = max(2, ntheta // 12), max(2, nphi // (4 * max(nports, 1)))
    for k in range(nports):
        jc = int((k + 0.5) * nphi / nports)
        for di in range(-pw_t // 2, pw_t // 2 + 1):
            for dj in range(-pw_p // 2, pw_p // 2 + 1):
                removed[di % ntheta, (jc + dj) % nphi] = True
    lc = []
    for i in range(ntheta):
        for j in range(nphi):
            if removed[i, j]:
                continue
            v00, v10, v01, v11 = vid(i, j), vid(i + 1, j), vid(i, j + 1), vid(i + 1, j + 1)
            if (i + j) % 2 == 0:
                lc += [[v00, v10, v11], [v00, v11, v01]]
            else:
                lc += [[v00, v10, v01], [v10, v11, v01]]
    lc = numpy.array(lc, dtype=numpy.int32)
    # drop vertices that lost all their cells (port interiors)
    used = numpy.zeros(r.shape[0], dtype=bool)
    used[lc.ravel()] = True
    new_id = -numpy.ones(r.shape[0], dtype=numpy.int64)
    new_id[used] = numpy.arange(used.sum())
    # hole loops: poloidal loop at the inboard-most phi column without ports, toroidal loop at theta = pi
    j_free = 0
    while removed[:, j_free].any() or removed[:, (j_free - 1) % nphi].any():
        j_free += 1
    pol = [vid(i, j_free) for i in range(ntheta)]
    i_in = ntheta // 2
    tor = [vid(i_in, j) for j in range(nphi)]
    nodesets = [new_id[numpy.array(pol)], new_id[numpy.array(tor)]]
    for k in range(nports):
        jc = int((k + 0.5) * nphi / nports)
        nodesets.append(new_id[numpy.array([vid(-pw_t // 2, jc)])])  # a vertex on the port rim (first removed row)
    r = r[used]
    lc = new_id[lc].astype(numpy.int32)
    closures = numpy.array([lc.shape[0] // 2], dtype=numpy.int32)
    if permute_seed is not None:
        perm = numpy.random.default_rng(permute_seed).permutation(r.shape[0])
        inv = numpy.empty_like(perm)
        inv[perm] = numpy.arange(r.shape[0])
        r = r[perm]
        lc = inv[lc].astype(numpy.int32)
        nodesets = [inv[ns] for ns in nodesets]
    return dict(r=numpy.ascontiguousarray(r), lc=numpy.ascontiguousarray(lc),
                nodesets=[numpy.asarray(ns, dtype=numpy.int32) for ns in nodesets], closures=closures)
