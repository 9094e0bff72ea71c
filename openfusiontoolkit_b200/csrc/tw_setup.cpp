// tw_setup.cpp -- O(N) host pre-processing of a ThinCurr model.
//
// What the reference does in tw_setup (src/physics/thin_wall.F90:166-523) and in the mesh
// layer it calls (src/grid/mesh_local.F90:105-267,809-1092; src/grid/trimesh_type.F90:243-249,
// 397-503,649-679), re-designed around flat hash/CSR containers and iterative traversals so it
// scales to 300k-triangle vessels.  The results that fix the numbering of the dense operators
// (edge ids, orientation flips, hole chains and signs, closure vertices, pmap) are defined so
// that they coincide with the reference's; tests compare them with the independent oracle.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <unordered_map>

#include "tw_host.h"

namespace tw {

// local edge slot -> local vertex pair (trimesh_type.F90:34, 0-based)
static const int kTriEd[3][2] = {{2, 1}, {0, 2}, {1, 0}};

static inline void cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

int32_t simple_hash(const void* key, long length) {
  // Jenkins one-at-a-time, as used for the cache-file model hashes (src/base/oft_local_c.c:86-98)
  const uint8_t* k = (const uint8_t*)key;
  uint32_t h = 0;
  for (long i = 0; i < length; i++) {
    h += k[i];
    h += h << 10;
    h ^= h >> 6;
  }
  h += h << 3;
  h ^= h >> 11;
  h += h << 15;
  return (int32_t)h;
}

int32_t Model::hash_lc() const {
  std::vector<int32_t> lc1(lc.size());
  for (size_t i = 0; i < lc.size(); i++) lc1[i] = lc[i] + 1;  // Fortran numbering in the reference's memory
  return simple_hash(lc1.data(), (long)(lc1.size() * 4));
}
int32_t Model::hash_r() const { return simple_hash(r.data(), (long)(r.size() * 8)); }

int Model::find_edge(int a, int b) const {
  int lo = std::min(a, b), hi = std::max(a, b);
  // edges of `lo` are stored contiguously (lexicographic numbering): binary search on hi
  // kpe lists all edges of a point; edges with this point as the low end form a sorted run.
  for (int k = kpe[lo]; k < kpe[lo + 1]; k++) {
    int e = lpe[k];
    if (le[2 * e] == lo && le[2 * e + 1] == hi) return e;
  }
  return -1;
}

void Model::invert_cell(int c) {
  // swap local vertices 2<->3 and the matching edge/neighbour slots (trimesh_type.F90:243-249)
  std::swap(lc[3 * c + 1], lc[3 * c + 2]);
  std::swap(lce[3 * c + 1], lce[3 * c + 2]);
  std::swap(lcc[3 * c + 1], lcc[3 * c + 2]);
}

std::string Model::mesh_init() {
  // ---- unique edges, numbered lexicographically by (lo,hi) (mesh_local.F90:105-205)
  std::vector<uint64_t> keys(3 * (size_t)nc);
  for (int c = 0; c < nc; c++)
    for (int j = 0; j < 3; j++) {
      int a = lc[3 * c + kTriEd[j][0]], b = lc[3 * c + kTriEd[j][1]];
      if (a < 0 || a >= np || b < 0 || b >= np) return "Cell list references a vertex outside the point list";
      if (a == b) return "Degenerate cell (repeated vertex)";
      keys[3 * (size_t)c + j] = ((uint64_t)std::min(a, b) << 32) | (uint32_t)std::max(a, b);
    }
  std::vector<uint64_t> uniq(keys);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  ne = (int)uniq.size();
  le.resize(2 * (size_t)ne);
  for (int e = 0; e < ne; e++) {
    le[2 * e] = (int)(uniq[e] >> 32);
    le[2 * e + 1] = (int)(uniq[e] & 0xffffffffu);
  }
  lce.resize(3 * (size_t)nc);
  for (size_t k = 0; k < keys.size(); k++)
    lce[k] = (int)(std::lower_bound(uniq.begin(), uniq.end(), keys[k]) - uniq.begin());
  // ---- point->cell and edge->cell CSR in ascending cell order (mesh_local.F90:213-267)
  auto build_csr = [](const std::vector<int>& item_of_slot, int nitems, int slots_per_cell, std::vector<int>& kp,
                      std::vector<int>& lp) {
    kp.assign(nitems + 1, 0);
    for (int v : item_of_slot) kp[v + 1]++;
    for (int i = 0; i < nitems; i++) kp[i + 1] += kp[i];
    lp.resize(item_of_slot.size());
    std::vector<int> fill(kp.begin(), kp.end() - 1);
    for (size_t s = 0; s < item_of_slot.size(); s++) lp[fill[item_of_slot[s]]++] = (int)(s / slots_per_cell);
  };
  build_csr(lc, np, 3, kpc, lpc);
  build_csr(lce, ne, 3, kec, lec);
  for (int v = 0; v < np; v++)
    if (kpc[v + 1] == kpc[v]) return "Floating vertex detected";
  // ---- neighbours across each local edge slot (mesh_local.F90:1022-1037)
  lcc.assign(3 * (size_t)nc, -1);
  for (int c = 0; c < nc; c++)
    for (int j = 0; j < 3; j++) {
      int e = lce[3 * c + j];
      int n = kec[e + 1] - kec[e];
      if (n == 2) lcc[3 * c + j] = lec[kec[e]] + lec[kec[e] + 1] - c;
      else if (n > 2) return "Non-manifold edge (more than two cells) is not supported";
    }
  if (!keep_orientation) sync_face_normals();
  // ---- boundary flags (mesh_local.F90:1044-1092)
  be.assign(ne, 0);
  bp.assign(np, 0);
  for (int e = 0; e < ne; e++)
    if (kec[e + 1] - kec[e] == 1) {
      be[e] = 1;
      bp[le[2 * e]] = bp[le[2 * e + 1]] = 1;
    }
  // ---- point->edge CSR, ascending edge id (mesh_local.F90:332-380)
  kpe.assign(np + 1, 0);
  for (int e = 0; e < ne; e++) {
    kpe[le[2 * e] + 1]++;
    kpe[le[2 * e + 1] + 1]++;
  }
  for (int i = 0; i < np; i++) kpe[i + 1] += kpe[i];
  lpe.resize(2 * (size_t)ne);
  std::vector<int> fill(kpe.begin(), kpe.end() - 1);
  for (int e = 0; e < ne; e++) {
    lpe[fill[le[2 * e]]++] = e;
    lpe[fill[le[2 * e + 1]]++] = e;
  }
  return "";
}

void Model::sync_face_normals() {
  // The reference orients each connected component from its lowest-numbered cell with a
  // recursive depth-first walk over neighbour slots 1..3 (mesh_local.F90:963-1014).  Same
  // visiting order here, with an explicit stack (recursion depth would reach nc).
  std::vector<char> done(nc, 0);
  std::vector<std::pair<int, int>> stack;
  nflipped = 0;
  for (int seed = 0; seed < nc; seed++) {
    if (done[seed]) continue;
    done[seed] = 1;
    stack.clear();
    stack.emplace_back(seed, 0);
    while (!stack.empty()) {
      int f1 = stack.back().first, j = stack.back().second;
      if (j == 3) {
        stack.pop_back();
        continue;
      }
      stack.back().second++;
      int f2 = lcc[3 * f1 + j];
      if (f2 < 0 || done[f2]) continue;
      int k = 0;
      while (k < 3 && lcc[3 * f2 + k] != f1) k++;
      // the shared edge must be traversed in opposite directions by consistently oriented cells
      bool same = lc[3 * f1 + kTriEd[j][0]] == lc[3 * f2 + kTriEd[k][0]] &&
                  lc[3 * f1 + kTriEd[j][1]] == lc[3 * f2 + kTriEd[k][1]];
      if (same) {
        invert_cell(f2);
        nflipped++;
      }
      done[f2] = 1;
      stack.emplace_back(f2, 0);
    }
  }
}

void Model::cell_normal(int c, double* n) const {
  // trimesh_tang / trimesh_norm for linear cells (trimesh_type.F90:649-679)
  const double *p0 = &r[3 * lc[3 * c]], *p1 = &r[3 * lc[3 * c + 1]], *p2 = &r[3 * lc[3 * c + 2]];
  double t1[3], t2[3];
  for (int d = 0; d < 3; d++) t1[d] = p1[d] - p0[d];
  double m = std::sqrt(dot(t1, t1));
  for (int d = 0; d < 3; d++) t1[d] = t1[d] / m;
  for (int d = 0; d < 3; d++) t2[d] = p2[d] - p0[d];
  double pr = dot(t2, t1);
  for (int d = 0; d < 3; d++) t2[d] = t2[d] - pr * t1[d];
  m = std::sqrt(dot(t2, t2));
  for (int d = 0; d < 3; d++) t2[d] = t2[d] / m;
  cross(t1, t2, n);
}

void Model::geometry() {
  // areas (mesh_local.F90:918-957), normals, grad(lambda) in the tangent frame
  // (trimesh_type.F90:397-503) and qbasis = grad(lambda_k) x n (thin_wall.F90:343-352)
  ca.resize(nc);
  va.assign(np, 0.0);
  norm.resize(3 * (size_t)nc);
  qbasis.resize(9 * (size_t)nc);
  for (int c = 0; c < nc; c++) {
    const double* p[3] = {&r[3 * lc[3 * c]], &r[3 * lc[3 * c + 1]], &r[3 * lc[3 * c + 2]]};
    double t1[3], t2[3], n[3];
    for (int d = 0; d < 3; d++) t1[d] = p[1][d] - p[0][d];
    double m = std::sqrt(dot(t1, t1));
    for (int d = 0; d < 3; d++) t1[d] = t1[d] / m;
    for (int d = 0; d < 3; d++) t2[d] = p[2][d] - p[0][d];
    double pr = dot(t2, t1);
    for (int d = 0; d < 3; d++) t2[d] = t2[d] - pr * t1[d];
    m = std::sqrt(dot(t2, t2));
    for (int d = 0; d < 3; d++) t2[d] = t2[d] / m;
    cross(t1, t2, n);
    double q[3][2];
    for (int k = 0; k < 3; k++) {
      q[k][0] = dot(p[k], t1);
      q[k][1] = dot(p[k], t2);
    }
    double A[2][2] = {{q[1][0] - q[0][0], q[1][1] - q[0][1]}, {q[2][0] - q[0][0], q[2][1] - q[0][1]}};
    double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    double C[2][2] = {{A[1][1] / det, -A[0][1] / det}, {-A[1][0] / det, A[0][0] / det}};
    double g[3][3];
    for (int d = 0; d < 3; d++) {
      g[1][d] = C[0][0] * t1[d] + C[1][0] * t2[d];
      g[2][d] = C[0][1] * t1[d] + C[1][1] * t2[d];
      g[0][d] = -(g[1][d] + g[2][d]);
    }
    ca[c] = std::fabs(det / 2.0);
    for (int d = 0; d < 3; d++) norm[3 * (size_t)c + d] = n[d];
    for (int k = 0; k < 3; k++) cross(g[k], n, &qbasis[9 * (size_t)c + 3 * k]);
  }
  for (int c = 0; c < nc; c++) {
    double a3 = ca[c] / 3.0;
    for (int k = 0; k < 3; k++) va[lc[3 * c + k]] += a3;
  }
}

// ------------------------------------------------------------------ holes
std::string Model::hole_pseq(int i0, std::vector<int>& chain) {
  // boundary loop through a seed vertex: always leave along the lowest-numbered boundary edge
  // that is not the one we arrived by (thin_wall.F90:405-440)
  if (i0 < 0 || i0 >= np || !bp[i0]) return "Hole starting vertex is not on boundary";
  int nbe = 0;
  for (char b : be) nbe += b;
  chain.assign(1, i0);
  int ipt = i0, eprev = -1;
  for (int step = 0; step < nbe; step++) {
    for (int k = kpe[ipt]; k < kpe[ipt + 1]; k++) {
      int ed = lpe[k];
      if (ed == eprev || !be[ed]) continue;
      ipt = le[2 * ed] + le[2 * ed + 1] - ipt;
      chain.push_back(ipt);
      eprev = ed;
      break;
    }
    if (ipt == i0) break;
  }
  if (ipt != i0) return "Error building hole mesh, could not find periodic path";
  chain.pop_back();
  return "";
}

std::string Model::order_hole_list(const std::vector<int>& in, std::vector<int>& out) {
  // Order an explicit vertex set into a closed chain (thin_wall.F90:444-522): greedy walk from
  // the lowest vertex that prefers successors which would not strand other set members
  // (a successor with >1 unvisited set-neighbours is only taken as a last resort).
  const int n = (int)in.size();
  std::vector<int> srt(in);
  std::sort(srt.begin(), srt.end());
  auto pos = [&](int v) -> int {
    auto it = std::lower_bound(srt.begin(), srt.end(), v);
    return (it != srt.end() && *it == v) ? (int)(it - srt.begin()) : -1;
  };
  std::vector<char> flag(n, 0);
  flag[0] = 1;
  int ipt = srt[0], eprev = -1;
  out.assign(1, ipt);
  for (int jj = 1; jj <= n; jj++) {
    if (jj == n - 2) flag[0] = 0;  // allow the walk to close on the first vertex
    int last_pt = -1, last_cand = -1, last_ed = -1;
    for (int k = kpe[ipt]; k < kpe[ipt + 1]; k++) {
      int ed = lpe[k];
      if (ed == eprev) continue;
      int ptp = le[2 * ed] + le[2 * ed + 1] - ipt;
      int cand = pos(ptp);
      if (cand < 0 || flag[cand]) continue;
      int nlinks = 0;
      for (int l = kpe[ptp]; l < kpe[ptp + 1]; l++) {
        int ed2 = lpe[l];
        int c2 = pos(le[2 * ed2] + le[2 * ed2 + 1] - ptp);
        if (c2 < 0 || flag[c2]) continue;
        nlinks++;
      }
      last_pt = ptp;
      last_cand = cand;
      last_ed = ed;
      if (nlinks > 1) continue;
      last_pt = -1;
      flag[cand] = 1;
      ipt = ptp;
      if (jj < n) {
        out.push_back(ipt);
        eprev = ed;
      }
      break;
    }
    if (last_pt >= 0) {
      flag[last_cand] = 1;
      ipt = last_pt;
      if (jj < n) {
        out.push_back(ipt);
        eprev = last_ed;
      }
    }
  }
  if ((int)out.size() != n) return "Error building hole mesh, unmatched points exist";
  if (ipt != srt[0]) return "Error building hole mesh, path is not periodic";
  return "";
}

std::string Model::setup_hole(const std::vector<int>& lp, std::vector<int>& cells_signed, std::vector<int>& kpc_h) {
  // Signed one-sided cell fan around an ordered closed vertex chain (thin_wall.F90:2230-2349).
  const int n = (int)lp.size();
  std::vector<int> fo(nc, 0), po(n, 0);
  for (int i = 0; i < n; i++) {
    int a = lp[i], b = lp[(i + 1) % n];
    int k = find_edge(a, b);
    if (k < 0) return "Could not find edge";
    double evec[3], ecc[3];
    for (int d = 0; d < 3; d++) {
      evec[d] = r[3 * b + d] - r[3 * a + d];
      ecc[d] = (r[3 * b + d] + r[3 * a + d]) / 2.0;
    }
    for (int j = kec[k]; j < kec[k + 1]; j++) {
      int c = lec[j];
      if (fo[c] != 0) continue;
      double ptcc[3], dv[3], cr[3], nn[3];
      for (int d = 0; d < 3; d++)
        ptcc[d] = (r[3 * lc[3 * c] + d] + r[3 * lc[3 * c + 1] + d] + r[3 * lc[3 * c + 2] + d]) / 3.0;
      cell_normal(c, nn);
      for (int d = 0; d < 3; d++) dv[d] = ptcc[d] - ecc[d];
      cross(dv, evec, cr);
      double val = dot(cr, nn);
      fo[c] = std::signbit(val) ? -1 : 1;
    }
    if (be[k]) {
      int s = fo[lec[kec[k]]];
      po[i] = s;
      po[(i + 1) % n] = s;
    }
  }
  bool all_nonneg = true, all_neg = true;
  for (int v : po) {
    all_nonneg &= (v >= 0);
    all_neg &= (v < 0);
  }
  if (all_nonneg) {
    std::fill(po.begin(), po.end(), 1);
  } else if (all_neg) {
    std::fill(po.begin(), po.end(), -1);
  } else {
    int prev = 0;
    for (int i = 0; i < n; i++) {
      if (po[i] == 0) {
        if (prev != 0) po[i] = prev;
      } else {
        prev = po[i];
      }
    }
    for (int i = 0; i < n; i++) {
      if (po[i] != 0) break;
      po[i] = prev;
    }
  }
  bool ok = true;
  for (int sweep = 0; sweep < 10; sweep++) {
    ok = true;
    for (int i = 0; i < n; i++)
      for (int j = kpc[lp[i]]; j < kpc[lp[i] + 1]; j++) {
        int c = lpc[j];
        if (fo[c] != 0) continue;
        for (int l = 0; l < 3; l++) {
          int f = lcc[3 * c + l];
          if (f < 0) continue;
          if (fo[f] != 0) {
            fo[c] = fo[f];
            break;
          }
        }
        if (fo[c] == 0) ok = false;
      }
    if (ok) break;
  }
  if (!ok) return "Error orienting cells";
  kpc_h.assign(1, 0);
  cells_signed.clear();
  for (int i = 0; i < n; i++) {
    for (int j = kpc[lp[i]]; j < kpc[lp[i] + 1]; j++) {
      int c = lpc[j];
      if (fo[c] == po[i]) cells_signed.push_back((c + 1) * fo[c]);
    }
    kpc_h.push_back((int)cells_signed.size());
  }
  return "";
}

std::string Model::build_holes(const std::vector<std::vector<int>>& nodesets0) {
  nholes = (int)nodesets0.size();
  hole_chain.assign(nholes, {});
  std::vector<std::vector<std::pair<int, int>>> per_cell(nc);
  std::vector<std::vector<int>> sorted_chains(nholes);
  for (int h = 0; h < nholes; h++) {
    const auto& ns = nodesets0[h];
    std::string err;
    if (ns.size() == 1)
      err = hole_pseq(ns[0], hole_chain[h]);
    else
      err = order_hole_list(ns, hole_chain[h]);
    if (!err.empty()) return err;
    std::vector<int> cells, kh;
    err = setup_hole(hole_chain[h], cells, kh);
    if (!err.empty()) return err;
    const auto& lp = hole_chain[h];
    for (size_t i = 0; i < lp.size(); i++)
      for (int k = kh[i]; k < kh[i + 1]; k++) {
        int c = std::abs(cells[k]) - 1, sg = cells[k] < 0 ? -1 : 1, l = 0;
        while (l < 3 && lc[3 * c + l] != lp[i]) l++;
        per_cell[c].emplace_back(sg * (h + 1), l);
      }
    sorted_chains[h] = lp;
    std::sort(sorted_chains[h].begin(), sorted_chains[h].end());
  }
  for (int i = 0; i < nholes; i++)
    for (int j = 0; j < i; j++)
      if (sorted_chains[i] == sorted_chains[j]) return "Duplicate hole detected";
  kfh.assign(nc + 1, 0);
  lfh.clear();
  for (int c = 0; c < nc; c++) {
    for (auto& e : per_cell[c]) {
      lfh.push_back(e.first);
      lfh.push_back(e.second);
    }
    kfh[c + 1] = (int)(lfh.size() / 2);
  }
  nfh = kfh[nc];
  return "";
}

std::string Model::build_pmap(const int* pmap_in, const std::vector<int>& closure_cells0) {
  // thin_wall.F90:282-320
  pmap.assign(np, 0);
  closures.clear();
  if (pmap_in == nullptr) {
    int k = 0;
    for (int v = 0; v < np; v++)
      if (!bp[v]) pmap[v] = ++k;
    for (int ci : closure_cells0) {
      if (ci < 0 || ci >= nc) return "Closure cell index out of range";
      int best = -1, j = 0;
      for (int kk = 0; kk < 3; kk++) {
        int v = lc[3 * ci + kk];
        if (pmap[v] <= 0) continue;
        int cnt = kpc[v + 1] - kpc[v];
        if (cnt > best) {
          best = cnt;
          j = kk;
        }
      }
      int v = lc[3 * ci + j];
      if (pmap[v] == 0) return "Error getting closure vertex";
      pmap[v] = -1;
      closures.push_back(v);
    }
    np_active = 0;
    for (int v = 0; v < np; v++) pmap[v] = (pmap[v] > 0) ? ++np_active : 0;
  } else {
    np_active = 0;
    for (int v = 0; v < np; v++) {
      pmap[v] = pmap_in[v];
      if (pmap[v] < 0) return "Invalid periodicity map";
      np_active = std::max(np_active, pmap[v]);
    }
  }
  return "";
}

std::string Model::setup_from_arrays(int np_, const double* r_, int nc_, const int* lc1, const int* reg_,
                                     const int* pmap_in, const std::vector<std::vector<int>>& nodesets0,
                                     const std::vector<int>& closure_cells0, const XmlNode* tc) {
  if (np_ < 3 || nc_ < 1) return "Mesh must contain at least one cell";
  np = np_;
  nc = nc_;
  r.assign(r_, r_ + 3 * (size_t)np);
  lc.resize(3 * (size_t)nc);
  for (size_t i = 0; i < lc.size(); i++) {
    if (lc1[i] < 1 || lc1[i] > np) return "Cell list refers to a vertex outside [1, np] (lc is 1-based on this interface)";
    lc[i] = lc1[i] - 1;
  }
  for (auto& ns : nodesets0)
    for (int v : ns)
      if (v < 0 || v >= np) return "Nodeset refers to a vertex outside the mesh";
  for (int c : closure_cells0)
    if (c < 0 || c >= nc) return "Closure (sideset 1) refers to a cell outside the mesh";
  reg.assign(nc, 1);
  if (reg_) reg.assign(reg_, reg_ + nc);
  nreg = 1;
  for (int v : reg) {
    if (v < 1) return "Region ids must be >= 1";
    nreg = std::max(nreg, v);
  }
  std::string err = mesh_init();
  if (!err.empty()) return err;
  // coils (thin_wall.F90:181-206)
  vcoils.clear();
  icoils.clear();
  if (tc) {
    if (const XmlNode* g = tc->child("vcoils")) {
      err = load_coils_xml(g, "VCOIL", vcoils);
      if (!err.empty()) return err;
    }
    for (auto& s : vcoils)
      for (auto& f : s.coils) {
        if (f.res_per_len < 0.0) return "Invalid resistivity for passive coil";
        if (f.radius < 1.e-6) return "Invalid radius for passive coil";
      }
    if (const XmlNode* g = tc->child("icoils")) {
      err = load_coils_xml(g, "ICOIL", icoils);
      if (!err.empty()) return err;
    }
    for (auto& s : icoils)
      for (auto& f : s.coils) f.radius = std::max(1.e-6, f.radius);
  }
  n_vcoils = (int)vcoils.size();
  n_icoils = (int)icoils.size();
  err = build_holes(nodesets0);
  if (!err.empty()) return err;
  err = build_pmap(pmap_in, closure_cells0);
  if (!err.empty()) return err;
  nelems = np_active + nholes + n_vcoils;
  geometry();
  eta_surf.assign(nreg, -1.0);
  eta_vol.assign(nreg, -1.0);
  thickness.assign(nreg, -1.0);
  sens_mask.assign(nreg, 0);
  if (tc) {
    err = load_eta_xml(tc);
    if (!err.empty()) return err;
  }
  return "";
}

// ------------------------------------------------------------------ resistance matrix
std::string Model::setup_from_tw(int np_, const double* r_, int nc_, const int* lc1, const int* reg_, const int* pmap1,
                                 int np_active_, int nholes_, const int* kfh1, const int* lfh1, const double* ca_,
                                 const double* qbasis_) {
  if (np_ < 3 || nc_ < 1) return "Mesh must contain at least one cell";
  np = np_;
  nc = nc_;
  r.assign(r_, r_ + 3 * (size_t)np);
  lc.resize(3 * (size_t)nc);
  for (size_t i = 0; i < lc.size(); i++) {
    if (lc1[i] < 1 || lc1[i] > np) return "Cell list refers to a vertex outside [1, np] (lc is 1-based on this interface)";
    lc[i] = lc1[i] - 1;
  }
  reg.assign(nc, 1);
  if (reg_) reg.assign(reg_, reg_ + nc);
  nreg = 1;
  for (int v : reg) nreg = std::max(nreg, v);
  keep_orientation = true;
  std::string err = mesh_init();
  if (!err.empty()) return err;
  vcoils.clear();
  icoils.clear();
  n_vcoils = n_icoils = 0;
  np_active = np_active_;
  nholes = nholes_;
  pmap.assign(pmap1, pmap1 + np);
  for (int v : pmap)
    if (v < 0 || v > np_active) return "pmap entry outside [0, np_active]";
  // Fortran CSR kfh(nc+1) (1-based offsets), lfh(2,nfh) column-major: (signed hole id, 1-based local vertex)
  kfh.resize(nc + 1);
  for (int c = 0; c <= nc; c++) kfh[c] = kfh1 ? kfh1[c] - 1 : 0;
  nfh = kfh[nc];
  lfh.resize(2 * (size_t)nfh);
  for (int i = 0; i < nfh; i++) {
    lfh[2 * i] = lfh1[2 * i];
    lfh[2 * i + 1] = lfh1[2 * i + 1] - 1;
    if (lfh[2 * i] == 0 || std::abs(lfh[2 * i]) > nholes || lfh[2 * i + 1] < 0 || lfh[2 * i + 1] > 2) return "Invalid hole incidence";
  }
  nelems = np_active + nholes + n_vcoils;
  geometry();
  // the host's own areas / basis vectors win so that both sides use bit-identical inputs
  if (ca_) ca.assign(ca_, ca_ + nc);
  if (qbasis_) qbasis.assign(qbasis_, qbasis_ + 9 * (size_t)nc);
  eta_surf.assign(nreg, -1.0);
  eta_vol.assign(nreg, -1.0);
  thickness.assign(nreg, -1.0);
  sens_mask.assign(nreg, 0);
  return "";
}

void Model::build_rmat() {
  // R[a][b] = sum_c eta_s(reg_c)/mu0 (E_c[a].E_c[b]) area_c, V-coil diagonal = R_coil/mu0
  // (thin_wall.F90:1690-1930).  Assembled through per-row ordered maps -> sorted 1-based CSR,
  // the structure the reference hands to Python (thincurr_f.F90:906-921).  Stays on the CPU.
  bool unset = true;
  for (double e : eta_surf) unset &= (e < 0.0);
  if (unset) std::fill(eta_surf.begin(), eta_surf.end(), 1.0);  // 'eta=mu0' fallback (:1704-1707)
  std::vector<std::map<int, double>> rows(nelems);
  std::vector<int> dof;
  std::vector<const double*> vec;
  std::vector<double> sgn;
  for (int c = 0; c < nc; c++) {
    dof.clear();
    vec.clear();
    sgn.clear();
    for (int k = 0; k < 3; k++) {
      int p = pmap[lc[3 * c + k]];
      if (p > 0) {
        dof.push_back(p - 1);
        vec.push_back(&qbasis[9 * (size_t)c + 3 * k]);
        sgn.push_back(1.0);
      }
    }
    for (int ii = kfh[c]; ii < kfh[c + 1]; ii++) {
      int h = lfh[2 * ii];
      dof.push_back(np_active + std::abs(h) - 1);
      vec.push_back(&qbasis[9 * (size_t)c + 3 * lfh[2 * ii + 1]]);
      sgn.push_back(h < 0 ? -1.0 : 1.0);
    }
    double eta = eta_surf[reg[c] - 1];
    for (size_t a = 0; a < dof.size(); a++)
      for (size_t b = 0; b < dof.size(); b++)
        rows[dof[a]][dof[b]] += eta * (sgn[a] * sgn[b] * dot(vec[a], vec[b])) * ca[c];
  }
  int ns = np_active + nholes;
  for (int i = 0; i < n_vcoils; i++) {
    double Rs = 0.0;
    for (auto& f : vcoils[i].coils) {
      double dl = 0.0;
      for (int k = 1; k < f.npts(); k++) {
        double d[3] = {f.pts[3 * k] - f.pts[3 * k - 3], f.pts[3 * k + 1] - f.pts[3 * k - 2], f.pts[3 * k + 2] - f.pts[3 * k - 1]};
        dl += std::sqrt(dot(d, d));
      }
      Rs += f.res_per_len * dl;
    }
    vcoils[i].Rself = Rs / kMu0;
    rows[ns + i][ns + i] += vcoils[i].Rself;
  }
  R_kr.assign(nelems + 1, 1);
  R_lc.clear();
  R_val.clear();
  for (int a = 0; a < nelems; a++) {
    for (auto& kv : rows[a]) {
      R_lc.push_back(kv.first + 1);
      R_val.push_back(kv.second);
    }
    R_kr[a + 1] = (int)R_lc.size() + 1;
  }
}

void FlatCoils::append(const CoilSet& s) {
  for (auto& f : s.coils) {
    pts.insert(pts.end(), f.pts.begin(), f.pts.end());
    fil_ptr.push_back(npts());
    scales.push_back(f.scale);
    radius.push_back(f.radius);
  }
  set_ptr.push_back(nfil());
  sens_mask.push_back(s.sens_mask ? 1 : 0);
}

}  // namespace tw
