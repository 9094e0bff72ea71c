// tw_plan.cpp -- patch / chunk / tile construction (see tw_plan.h).
#include "tw_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

namespace tw {

namespace {
// kd-order of points: recursive split along the longest bounding-box axis.  `npatch` leaves of (nearly) equal
// size become patches (the split point is proportional to the leaf counts of the two halves, so any patch count
// works, not only powers of two); recursion continues below a leaf only to order the DOFs inside the patch.
void rcb(std::vector<int>& idx, int lo, int hi, const std::vector<double>& xyz, int npatch, std::vector<int>& cuts) {
  int n = hi - lo;
  if (npatch == 1) cuts.push_back(lo);
  if (n <= 4) return;
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = lo; i < hi; i++)
    for (int d = 0; d < 3; d++) {
      double v = xyz[3 * (size_t)idx[i] + d];
      mn[d] = std::min(mn[d], v);
      mx[d] = std::max(mx[d], v);
    }
  int ax = 0;
  for (int d = 1; d < 3; d++)
    if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
  const int left = npatch > 1 ? npatch / 2 : 0;
  int mid = npatch > 1 ? lo + (int)((long long)n * left / npatch) : lo + n / 2;
  std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
    double va = xyz[3 * (size_t)a + ax], vb = xyz[3 * (size_t)b + ax];
    return va < vb || (va == vb && a < b);
  });
  rcb(idx, lo, mid, xyz, npatch > 1 ? left : 0, cuts);
  rcb(idx, mid, hi, xyz, npatch > 1 ? npatch - left : 0, cuts);
}
}  // namespace

int auto_patch_size(int nv, int nshards) {
  int P;
  {
    // Large patches keep the halo (cells shared by neighbouring patches are evaluated once per patch) small, but the tile
    // count must fill the 148 SMs of every device many times over (>= ~1750 tiles per device of the upper triangle):
    // P = nv / (59 sqrt(nshards)), between 300 and 1200 DOFs.  Measured with the round-2 kernel on the 100k-vertex vessel,
    // one device: 526 -> 1732 ms, 800 -> 1659 ms, 1200 -> 1655 ms, 2000 -> 1670 ms; 20k vessel (round 1): 150 -> 186 ms,
    // 300 -> 169 ms, 600 -> 235 ms.  Small meshes shrink the patches until there are ~1500 tiles (>= 32 DOFs).
    // Meshes below ~60k vertices get the same patches for every shard count, i.e. the same bits.
    P = std::min(1200, std::max(300, (int)(nv / (59.0 * std::sqrt((double)std::max(1, nshards))))));
    while (P > 32 && ((long)((nv + P - 1) / P) * ((nv + P - 1) / P)) / 2 < 1500) P = P * 3 / 4;
    P = std::max(P, 32);
  }
  return P;
}

std::string build_patches(const Model& m, int P, PatchSet& ps, int nshards, const std::vector<int>* ref_cuts, std::vector<int>* band_patch_ptr) {
  const int nv = m.np_active, nh = m.nholes;
  ps = PatchSet();
  ps.ndof = nv + nh;
  if (P <= 0) P = auto_patch_size(nv, nshards);
  // dof -> vertices (periodic meshes map several vertices to one DOF)
  std::vector<int> kdv(nv + 1, 0), ldv;
  for (int v = 0; v < m.np; v++)
    if (m.pmap[v] > 0) kdv[m.pmap[v]]++;
  for (int i = 0; i < nv; i++) kdv[i + 1] += kdv[i];
  ldv.resize(kdv[nv]);
  {
    std::vector<int> fill(kdv.begin(), kdv.end() - 1);
    for (int v = 0; v < m.np; v++)
      if (m.pmap[v] > 0) ldv[fill[m.pmap[v] - 1]++] = v;
  }
  for (int d = 0; d < nv; d++)
    if (kdv[d + 1] == kdv[d]) return "Invalid periodicity map (unused DOF id)";
  std::vector<double> xyz(3 * (size_t)std::max(nv, 1));
  for (int d = 0; d < nv; d++)
    for (int k = 0; k < 3; k++) xyz[3 * (size_t)d + k] = m.r[3 * (size_t)ldv[kdv[d]] + k];
  std::vector<int> idx(nv), cuts;
  std::iota(idx.begin(), idx.end(), 0);
  if (ref_cuts && ref_cuts->size() >= 2 && ref_cuts->front() == 0 && ref_cuts->back() == nv) {
    // patches per range of reference ids (idx is the identity here: a range of idx is a range of reference ids)
    for (size_t b = 0; b + 1 < ref_cuts->size(); b++) {
      const int lo = (*ref_cuts)[b], hi = (*ref_cuts)[b + 1];
      if (band_patch_ptr) band_patch_ptr->push_back((int)cuts.size());
      if (hi > lo) rcb(idx, lo, hi, xyz, std::max(1, (hi - lo + P / 2) / P), cuts);
    }
    if (band_patch_ptr) band_patch_ptr->push_back((int)cuts.size());
  } else if (nv > 0) {
    rcb(idx, 0, nv, xyz, std::max(1, (nv + P - 1) / P), cuts);
  }
  cuts.push_back(nv);
  ps.nvert_patch = (int)cuts.size() - 1;
  ps.npatch = ps.nvert_patch + nh;
  ps.dof_orig.resize(ps.ndof);
  ps.patch_dof_ptr.assign(1, 0);
  // Internal numbering inside a patch = ascending reference id (rows of a patch leave the device as a few long runs of
  // consecutive reference rows); the kd order `idx` is kept for the traversal that forms the (compact) chunks.
  std::vector<int> loc(std::max(nv, 1), 0);  // reference DOF -> internal index
  for (int p = 0; p < ps.nvert_patch; p++) {
    for (int i = cuts[p]; i < cuts[p + 1]; i++) ps.dof_orig[i] = idx[i];
    std::sort(ps.dof_orig.begin() + cuts[p], ps.dof_orig.begin() + cuts[p + 1]);
    for (int i = cuts[p]; i < cuts[p + 1]; i++) loc[ps.dof_orig[i]] = i;
    ps.patch_dof_ptr.push_back(cuts[p + 1]);
  }
  for (int h = 0; h < nh; h++) {
    ps.dof_orig[nv + h] = nv + h;
    ps.patch_dof_ptr.push_back(nv + h + 1);
  }
  // hole incidences grouped per hole, in cell order
  std::vector<std::vector<std::pair<int, int>>> hole_inc(nh);  // (cell, k | neg<<2)
  for (int c = 0; c < m.nc; c++)
    for (int ii = m.kfh[c]; ii < m.kfh[c + 1]; ii++) {
      int h = m.lfh[2 * ii];
      hole_inc[std::abs(h) - 1].emplace_back(c, m.lfh[2 * ii + 1] | (h < 0 ? 4 : 0));
    }
  ps.patch_chunk_ptr.assign(1, 0);
  ps.patch_ncell.clear();
  std::vector<int> cell_slot(m.nc, -1);
  struct Inc {
    int dof_local, cell_slot, code;
  };
  std::vector<Inc> incs;
  std::vector<int> cells;
  for (int p = 0; p < ps.npatch; p++) {
    incs.clear();
    cells.clear();
    int d0 = ps.patch_dof_ptr[p], d1 = ps.patch_dof_ptr[p + 1];
    auto add = [&](int dl, int c, int code) {
      if (cell_slot[c] < 0) {
        cell_slot[c] = (int)cells.size();
        cells.push_back(c);
      }
      incs.push_back({dl, cell_slot[c], code});
    };
    if (p < ps.nvert_patch) {
      for (int ti = d0; ti < d1; ti++) {
        int d = idx[ti];
        for (int kv = kdv[d]; kv < kdv[d + 1]; kv++) {
          int v = ldv[kv];
          for (int j = m.kpc[v]; j < m.kpc[v + 1]; j++) {
            int c = m.lpc[j], k = 0;
            while (k < 3 && m.lc[3 * c + k] != v) k++;
            add(loc[d] - d0, c, k);
          }
        }
      }
    } else {
      for (auto& e : hole_inc[p - ps.nvert_patch]) add(0, e.first, e.second);
    }
    int ncell = (int)cells.size();
    ps.patch_ncell.push_back(ncell);
    int nch = (ncell + kCH - 1) / kCH;
    // bucket incidences by chunk, keeping (dof, cell) order
    std::vector<std::vector<Inc>> by_chunk(nch);
    for (auto& e : incs) by_chunk[e.cell_slot / kCH].push_back(e);
    for (int ch = 0; ch < nch; ch++) {
      int chunk_id = (int)ps.chunks.size();
      ChunkMeta cm;
      cm.ncell = std::min(kCH, ncell - ch * kCH);
      cm.dof_off = (int)ps.chunk_dof.size();
      cm.inc_off = (int)ps.inc.size();
      size_t g0 = ps.geom.size();
      ps.geom.resize(g0 + (size_t)kGeomRows * kCH, 0.0);
      ps.cell_dmin.resize(ps.cell_dmin.size() + kCH, 0x7fffffff);
      ps.cell_dmax.resize(ps.cell_dmax.size() + kCH, -1);
      ps.cell_ids.resize(ps.cell_ids.size() + kCH, -1);
      for (int s = 0; s < cm.ncell; s++) {
        int c = cells[ch * kCH + s];
        ps.cell_ids[(size_t)chunk_id * kCH + s] = c;
        for (int k = 0; k < 3; k++)
          for (int d = 0; d < 3; d++) {
            ps.geom[g0 + (size_t)(k * 3 + d) * kCH + s] = m.r[3 * (size_t)m.lc[3 * c + k] + d];
            ps.geom[g0 + (size_t)(10 + k * 3 + d) * kCH + s] = m.qbasis[9 * (size_t)c + 3 * k + d];
          }
        ps.geom[g0 + (size_t)9 * kCH + s] = m.ca[c];
        double P9[9], nh[3];
        for (int k = 0; k < 3; k++)
          for (int d = 0; d < 3; d++) P9[3 * k + d] = m.r[3 * (size_t)m.lc[3 * c + k] + d];
        phipot_normal(P9, nh);
        for (int d = 0; d < 3; d++) {
          ps.geom[g0 + (size_t)(19 + d) * kCH + s] = nh[d];
          ps.geom[g0 + (size_t)(22 + d) * kCH + s] = m.norm[3 * (size_t)c + d];
        }
      }
      {
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int s = 0; s < cm.ncell; s++)
          for (int k = 0; k < 3; k++)
            for (int d = 0; d < 3; d++) {
              double v = ps.geom[g0 + (size_t)(k * 3 + d) * kCH + s];
              mn[d] = std::min(mn[d], v);
              mx[d] = std::max(mx[d], v);
            }
        cm.cx = 0.5 * (mn[0] + mx[0]);
        cm.cy = 0.5 * (mn[1] + mx[1]);
        cm.cz = 0.5 * (mn[2] + mx[2]);
        double r2 = 0.0;
        for (int s = 0; s < cm.ncell; s++)
          for (int k = 0; k < 3; k++) {
            double dx = ps.geom[g0 + (size_t)(k * 3) * kCH + s] - cm.cx, dy = ps.geom[g0 + (size_t)(k * 3 + 1) * kCH + s] - cm.cy,
                   dz = ps.geom[g0 + (size_t)(k * 3 + 2) * kCH + s] - cm.cz;
            r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
          }
        cm.rad = std::sqrt(r2) * (1.0 + 1e-12);
        double e2 = 0.0;
        for (int s = 0; s < cm.ncell; s++)
          for (int k = 0; k < 3; k++) {
            const int k1 = (k + 1) % 3;
            double dx = ps.geom[g0 + (size_t)(k * 3) * kCH + s] - ps.geom[g0 + (size_t)(k1 * 3) * kCH + s],
                   dy = ps.geom[g0 + (size_t)(k * 3 + 1) * kCH + s] - ps.geom[g0 + (size_t)(k1 * 3 + 1) * kCH + s],
                   dz = ps.geom[g0 + (size_t)(k * 3 + 2) * kCH + s] - ps.geom[g0 + (size_t)(k1 * 3 + 2) * kCH + s];
            e2 = std::max(e2, dx * dx + dy * dy + dz * dz);
          }
        cm.emax = std::sqrt(e2) * (1.0 + 1e-12);
      }
      auto& L = by_chunk[ch];
      // local DOFs in ascending reference id: neighbouring lanes of the kernel's write-out touch neighbouring columns
      std::stable_sort(L.begin(), L.end(), [&](const Inc& a, const Inc& b) { return ps.dof_orig[d0 + a.dof_local] < ps.dof_orig[d0 + b.dof_local]; });
      std::vector<int> ptr;
      int prev = -1;
      for (size_t k = 0; k < L.size(); k++) {
        if (L[k].dof_local != prev) {
          ps.chunk_dof.push_back(d0 + L[k].dof_local);
          ptr.push_back((int)k);
          prev = L[k].dof_local;
        }
        int s = L[k].cell_slot - ch * kCH;
        ps.inc.push_back((uint16_t)(s | ((L[k].code & 3) << 6) | ((L[k].code & 4) ? 256 : 0)));
        int od = ps.dof_orig[d0 + L[k].dof_local];
        int& mn = ps.cell_dmin[(size_t)chunk_id * kCH + s];
        int& mx = ps.cell_dmax[(size_t)chunk_id * kCH + s];
        mn = std::min(mn, od);
        mx = std::max(mx, od);
      }
      ptr.push_back((int)L.size());
      cm.ndof = (int)ptr.size() - 1;
      if (cm.ndof > kMaxChunkDof || (int)L.size() > kMaxChunkInc) return "Internal error: chunk incidence overflow";
      // inc_ptr for chunk i lives at [dof_off + chunk_id, dof_off + chunk_id + ndof]
      for (int v : ptr) ps.chunk_inc_ptr.push_back(v);
      ps.chunks.push_back(cm);
    }
    ps.patch_chunk_ptr.push_back((int)ps.chunks.size());
    for (int c : cells) cell_slot[c] = -1;
  }
  ps.nchunk = (int)ps.chunks.size();
  return "";
}

void shard_range(const PatchSet& ps, int nshards, int shard, int& p0, int& p1) {
  // contiguous patch ranges with (nearly) equal numbers of row cells: the work of a row patch
  // is ncell(patch) x (all column cells)
  double total = 0;
  for (int n : ps.patch_ncell) total += n;
  auto cut = [&](int s) {
    if (s <= 0) return 0;
    if (s >= nshards) return ps.npatch;
    double target = total * s / nshards, acc = 0;
    for (int p = 0; p < ps.npatch; p++) {
      if (acc + 0.5 * ps.patch_ncell[p] >= target) return p;
      acc += ps.patch_ncell[p];
    }
    return ps.npatch;
  };
  p0 = cut(shard);
  p1 = cut(shard + 1);
}

namespace {
struct Ball {
  double c[3], r;
};
std::vector<Ball> patch_balls(const PatchSet& ps) {
  std::vector<Ball> out(ps.npatch);
  for (int p = 0; p < ps.npatch; p++) {
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int ch = ps.patch_chunk_ptr[p]; ch < ps.patch_chunk_ptr[p + 1]; ch++)
      for (int s = 0; s < ps.chunks[ch].ncell; s++)
        for (int k = 0; k < 3; k++)
          for (int d = 0; d < 3; d++) {
            double v = ps.geom[((size_t)ch * kGeomRows + k * 3 + d) * kCH + s];
            mn[d] = std::min(mn[d], v);
            mx[d] = std::max(mx[d], v);
          }
    Ball b;
    double r2 = 0;
    for (int d = 0; d < 3; d++) {
      b.c[d] = 0.5 * (mn[d] + mx[d]);
      r2 += 0.25 * (mx[d] - mn[d]) * (mx[d] - mn[d]);
    }
    b.r = std::sqrt(r2);
    out[p] = b;
  }
  return out;
}
double mean_cell_size(const PatchSet& ps) {
  double a = 0;
  long n = 0;
  for (int ch = 0; ch < ps.nchunk; ch++)
    for (int s = 0; s < ps.chunks[ch].ncell; s++) {
      a += ps.geom[((size_t)ch * kGeomRows + 9) * kCH + s];
      n++;
    }
  return n ? std::sqrt(2.0 * a / n) : 1.0;
}
// Relative cost of one cell pair by the distance D of its chunks' centres in mean cell sizes h: the quadrature order falls
// with D (thin_wall.F90:1055-1059: order >= k while 1 - dl_min/dl_max >= exp(ln(1e-8)/k), and dl_max - dl_min is about one
// cell size), the cost of a far pair is its n^2 evaluations plus ~25 evaluation-equivalents of classification and
// contraction, a near pair costs 28..72 analytic potentials of ~50 evaluation-equivalents each.  Normalised to order 4.
double pair_weight(double D_over_h) {
  const double x = D_over_h;
  if (x > 39.8) return 1.0;     // order 4 (6 points)
  if (x > 21.5) return 1.21;    // 5 (7)
  if (x > 13.9) return 2.8;     // 6 (12)
  if (x > 10.0) return 4.1;     // 7 (15)
  if (x > 7.7) return 4.6;      // 8 (16)
  if (x > 6.3) return 6.3;      // 9 (19)
  if (x > 5.3) return 10.7;     // 10 (25)
  return 42.0;                  // near field
}
float tile_cost(const PatchSet& A, const PatchSet& B, const std::vector<Ball>& ba, const std::vector<Ball>& bb, int pa,
                int pb, double h) {
  // chunk pair by chunk pair when the patches are close (the mix of orders varies across the tile), one weight otherwise
  double d = 0;
  for (int k = 0; k < 3; k++) d += (ba[pa].c[k] - bb[pb].c[k]) * (ba[pa].c[k] - bb[pb].c[k]);
  d = std::sqrt(d);
  if (d - ba[pa].r - bb[pb].r > 39.8 * h) return (float)((double)A.patch_ncell[pa] * B.patch_ncell[pb]);
  double cost = 0.0;
  for (int ci = A.patch_chunk_ptr[pa]; ci < A.patch_chunk_ptr[pa + 1]; ci++) {
    const ChunkMeta& I = A.chunks[ci];
    for (int cj = B.patch_chunk_ptr[pb]; cj < B.patch_chunk_ptr[pb + 1]; cj++) {
      const ChunkMeta& J = B.chunks[cj];
      const double dx = I.cx - J.cx, dy = I.cy - J.cy, dz = I.cz - J.cz;
      cost += (double)I.ncell * J.ncell * pair_weight(std::sqrt(dx * dx + dy * dy + dz * dz) / h);
    }
  }
  return (float)cost;
}
// cost estimates of all self tiles of a patch set (symmetric), computed once per plan
const std::vector<float>& self_costs(const PatchSet& ps) {
  if (ps.self_cost.size() == (size_t)ps.npatch * ps.npatch) return ps.self_cost;
  auto balls = patch_balls(ps);
  const double h = mean_cell_size(ps);
  ps.self_cost.assign((size_t)ps.npatch * ps.npatch, 0.f);
  for (int p = 0; p < ps.npatch; p++)
    for (int q = p; q < ps.npatch; q++) {
      const float c = tile_cost(ps, ps, balls, balls, p, q, h);
      ps.self_cost[(size_t)p * ps.npatch + q] = c;
      ps.self_cost[(size_t)q * ps.npatch + p] = c;
    }
  return ps.self_cost;
}
}  // namespace

#if defined(__GNUC__)
__attribute__((optimize("fp-contract=off")))
#endif
void phipot_normal(const double* P, double* n) {
  volatile double a0 = P[3] - P[0], a1 = P[4] - P[1], a2 = P[5] - P[2];
  volatile double b0 = P[6] - P[3], b1 = P[7] - P[4], b2 = P[8] - P[5];
  volatile double t0 = a1 * b2, t1 = a2 * b1, t2 = a2 * b0, t3 = a0 * b2, t4 = a0 * b1, t5 = a1 * b0;
  volatile double n0 = t0 - t1, n1 = t2 - t3, n2 = t4 - t5;
  volatile double s0 = n0 * n0, s1 = n1 * n1, s2 = n2 * n2;
  volatile double s01 = s0 + s1;
  volatile double ss = s01 + s2;
  const double m = std::sqrt(ss);
  n[0] = n0 / m;
  n[1] = n1 / m;
  n[2] = n2 / m;
}

void shard_range_sym(const PatchSet& ps, int nshards, int shard, int& p0, int& p1) {
  // Every tile {p,q} between two shards is evaluated by exactly one of them (sym_tile_is_mine: a checkerboard, so each
  // off-diagonal block is split evenly between its two shards), tiles inside a shard by that shard.  The work attributable
  // to row patch p is therefore half the cost of its whole row of tiles (same cost model as the tile queue: near-field
  // tiles weigh more); contiguous patch ranges with equal sums.
  const std::vector<float>& cost = self_costs(ps);
  std::vector<double> w(ps.npatch, 0.0);
  double total = 0.0;
  for (int p = 0; p < ps.npatch; p++) {
    for (int q = 0; q < ps.npatch; q++) w[p] += 0.5 * (double)cost[(size_t)p * ps.npatch + q];
    total += w[p];
  }
  auto cut = [&](int s) {
    if (s <= 0) return 0;
    if (s >= nshards) return ps.npatch;
    double target = total * s / nshards, acc = 0;
    for (int p = 0; p < ps.npatch; p++) {
      if (acc + 0.5 * w[p] >= target) return p;
      acc += w[p];
    }
    return ps.npatch;
  };
  p0 = cut(shard);
  p1 = cut(shard + 1);
}

void build_self_tiles(const PatchSet& ps, int p0, int p1, std::vector<Tile>& tiles, bool upper_only, int skip_lo) {
  tiles.clear();
  const std::vector<float>& cost = self_costs(ps);
  std::vector<int> omin(ps.npatch, 0x7fffffff), omax(ps.npatch, -1);
  for (int p = 0; p < ps.npatch; p++)
    for (int i = ps.patch_dof_ptr[p]; i < ps.patch_dof_ptr[p + 1]; i++) {
      omin[p] = std::min(omin[p], ps.dof_orig[i]);
      omax[p] = std::max(omax[p], ps.dof_orig[i]);
    }
  for (int pa = p0; pa < p1; pa++)
    for (int pb = 0; pb < ps.npatch; pb++) {
      bool owned = pb >= p0 && pb < p1;
      if (owned && pb < pa) continue;  // produced by the mirror write of tile (pb,pa)
      if (skip_lo >= 0 && pb >= skip_lo && pb < p0) continue;  // rows of an earlier band of the same device (copied afterwards)
      Tile t;
      t.flags = 0;
      if (upper_only && !owned) {
        // symmetric shards: a tile between two shards belongs to one of them and is evaluated with the owner's rows as the
        // row side (direct, coalesced writes); the other shard copies the transposed block afterwards (exchange)
        if (!sym_tile_is_mine(pa, pb)) continue;
        t.pa = pa;
        t.pb = pb;
      } else if (!owned && pb < pa) {
        // rows of pa against an unowned lower patch: evaluate the tile in the orientation (pb,pa) the
        // single-device build uses and keep only its mirror writes, so that every shard count
        // produces the same bits (row_out is -1 on the unowned side: no direct write happens)
        t.pa = pb;
        t.pb = pa;
        t.flags |= 2;
      } else {
        t.pa = pa;
        t.pb = pb;
        if (pa == pb) t.flags |= 1 | 8;
        else if (owned) t.flags |= 2 | 8;  // both row blocks owned: the transposed entries are filled by the symmetrisation pass
      }
      if (t.pa != t.pb && omax[t.pa] > omin[t.pb]) t.flags |= 4;
      t.cost = cost[(size_t)t.pa * ps.npatch + t.pb] * ((t.flags & 1) ? 0.5f : 1.0f);
      tiles.push_back(t);
    }
  std::stable_sort(tiles.begin(), tiles.end(), [](const Tile& a, const Tile& b) { return a.cost > b.cost; });
}

void build_mutual_tiles(const PatchSet& rows, const PatchSet& cols, std::vector<Tile>& tiles) {
  tiles.clear();
  auto ba = patch_balls(rows), bb = patch_balls(cols);
  double h = std::max(mean_cell_size(rows), mean_cell_size(cols));
  for (int pa = 0; pa < rows.npatch; pa++)
    for (int pb = 0; pb < cols.npatch; pb++) {
      Tile t{pa, pb, 0, tile_cost(rows, cols, ba, bb, pa, pb, h)};
      tiles.push_back(t);
    }
  std::stable_sort(tiles.begin(), tiles.end(), [](const Tile& a, const Tile& b) { return a.cost > b.cost; });
}

}  // namespace tw
