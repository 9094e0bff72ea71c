// tw_blocks.cu -- pair-integral sweeps over cell LISTS:
//
//  * tw_compute_Lmatblock / tw_compute_LmatHole / tw_compute_Bops_block (src/physics/thin_wall_hodlr.F90:136-404,
//    580-691): the dense-block evaluators the HODLR/ACA+ compression calls -- blocks of a few hundred DOFs and strips
//    of ONE row DOF (its ~6 cells) against a whole column block (:1224-1428).  Unlike tw_compute_LmatDirect there is
//    no role rule and no skipped pair: the ROW block's cell is always the analytic side of a near pair.
//  * tw_compute_Lmat_MF (src/physics/thin_wall.F90:1190-1414): matrix-free b = M a between two models with its own
//    3-level quadrature heuristic (behind ThinCurr.cross_eval, thincurr_f.F90:525-541).
//  * the projections of tw_reduce_model (src/physics/thin_wall_solvers.F90:1180-1359).
//
// One sweep skeleton serves all of them: a CTA owns 256 column items (cells, or mesh vertices for the B operator), one
// per thread, and walks the row cells in groups of 16 staged in shared memory together with the points of ALL rules
// 4..10 (100 points, built once per group and reused by the 256 columns).  Far pairs are evaluated in place by the
// owning thread (rule chosen per pair, bit-identical to the reference's decision); near pairs go to a per-warp list
// (ballot order) and are evaluated by the whole warp, lanes over the evaluation points of the analytic potential.
// What happens to T(i,j) differs: stored (dense T[rows][cols] in HBM, contracted afterwards by a gather kernel in a
// fixed order: deterministic, no atomics) or consumed at once (matrix-free apply: F_j += T(i,j) J_i).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/thincurr_b200.h"
#include "tw_gpu.h"
#include "tw_ops.h"

namespace twk {

constexpr int kSwT = 256;       // threads = column items per CTA
constexpr int kSwR = 16;        // row cells per staged group
constexpr int kSwPts = 100;    // points of the rules 4..10 (TCQ_OFF[4] = 7 .. TCQ_OFF[11] = 107)
constexpr int kSwOff = 7;
constexpr int kMfQ = 4;         // right-hand sides per matrix-free sweep

struct RowSlot {
  double P[9];
  double area;
  double nh[3];    // unit normal as tw_compute_phipot forms it
  double nrm[3];   // mesh normal (B operator: on-surface offset direction)
  double pts[kSwPts * 4];  // x, y, z, weight
};

struct SweepSmem {
  RowSlot row[kSwR];
  double J[kSwR][kMfQ][3];             // matrix-free: current vector of the row cells for the rhs group
  uint16_t list[kSwT / 32][kSwR * 32];  // near pairs per warp: slot << 5 | lane
  unsigned int blist[kSwR * kSwT];      // B operator: block-wide near list  slot << 16 | neighbor << 15 | thread
  int bcount;
};

struct SweepArgs {
  const double *Pr, *Ar, *Nr;  // row model: cell vertices [nc][9], areas, mesh normals [nc][3]
  const double *Pc, *Ac;       // column model cells (mode 0/1) ...
  const double *rc, *vac;      // ... or column model vertices [np][3] and vertex areas (mode 2)
  const int* row_cells;        // [nrc] cell ids or NULL (identity)
  const int* col_items;        // [ncc] cell / vertex ids or NULL (identity)
  int nrc, ncc;
  int row0, row1;              // row range of this launch (indices into row_cells); T rows are relative to row0
  int rows_per_y;
  double* T;                   // mode 0: [row1-row0][ldT];  mode 2: D[row1-row0][ldT][3]
  long long ldT;
  const double* J;             // mode 1: [nrc][kMfQ][3]
  double* F;                   // mode 1: [gridDim.y][ncc][kMfQ][3]
  unsigned long long* counts;  // optional [3]: pairs per class (mode 1) / far, near (mode 0, 2)
  // modes 0 / 2: near pairs are not evaluated by the sweep but appended here (row-list index, column index) and handed
  // to a second kernel that spreads them over the whole device (they cluster in the few CTAs whose columns touch the
  // staged rows); beyond near_cap the sweep evaluates them itself
  int2* near_list;
  unsigned int* near_count;
  unsigned int near_cap;
};

// ---- classification of tw_compute_Lmat_MF (thin_wall.F90:1243-1288) --------------------------------------------
// 0 far (rule 6), 1 close (rule 10), 2 very close (analytic).  Angles are compared through their cosines and distance
// ratios through their squares; inside a guard band of a threshold the reference expression itself is evaluated.
__device__ __noinline__ int mf_class_exact(const double* Pi, const double* Pj) {
  bool close_flag = false, vv = false;
  double dl_max = -1.e99;
  for (int ii = 0; ii < 3; ii++) {
    double a[3], b[3];
    for (int d = 0; d < 3; d++) a[d] = xsub(Pj[d], Pi[3 * ii + d]);
    double t = __dsqrt_rn(xdot(a[0], a[1], a[2], a[0], a[1], a[2]));
    if (t < 1.e-10) { close_flag = true; break; }
    for (int d = 0; d < 3; d++) a[d] = __ddiv_rn(a[d], t);
    for (int d = 0; d < 3; d++) b[d] = xsub(Pj[3 + d], Pi[3 * ii + d]);
    t = __dsqrt_rn(xdot(b[0], b[1], b[2], b[0], b[1], b[2]));
    if (t < 1.e-10) { close_flag = true; break; }
    for (int d = 0; d < 3; d++) b[d] = __ddiv_rn(b[d], t);
    dl_max = fmax(dl_max, fabs(acos(xdot(a[0], a[1], a[2], b[0], b[1], b[2]))));
    for (int d = 0; d < 3; d++) b[d] = xsub(Pj[6 + d], Pi[3 * ii + d]);
    t = __dsqrt_rn(xdot(b[0], b[1], b[2], b[0], b[1], b[2]));
    if (t < 1.e-10) { close_flag = true; break; }
    for (int d = 0; d < 3; d++) b[d] = __ddiv_rn(b[d], t);
    dl_max = fmax(dl_max, fabs(acos(xdot(a[0], a[1], a[2], b[0], b[1], b[2]))));
  }
  const double pi = 3.14159265358979323846;
  if (dl_max > pi / 8.0) {
    close_flag = true;
    if (dl_max > pi / 4.0) vv = true;
  } else {
    double dmin = 1.e99, dmax = -1.e99;
    for (int ii = 0; ii < 3; ii++)
      for (int jj = 0; jj < 3; jj++) {
        double dx = xsub(Pi[3 * ii], Pj[3 * jj]), dy = xsub(Pi[3 * ii + 1], Pj[3 * jj + 1]), dz = xsub(Pi[3 * ii + 2], Pj[3 * jj + 2]);
        double d = __dsqrt_rn(xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz)));
        dmin = fmin(dmin, d);
        dmax = fmax(dmax, d);
      }
    const double rho = __ddiv_rn(dmin, dmax);
    if (rho < 0.95) close_flag = true;
    if (rho < 0.75) vv = true;
  }
  return close_flag ? (vv ? 2 : 1) : 0;
}

__device__ __forceinline__ int mf_class(const double* Pi, const double* Pj) {
  const double cA = 0.92387953251128674, cB = 0.70710678118654752;  // cos(pi/8), cos(pi/4)
  const double band = 1.e-11;
  double cmin = 2.0;  // min cosine = max angle (2: no angle formed yet)
  bool hit = false;   // a coincident vertex ended the angle loop
  double d2min = 1.e300, d2max = 0.0;
#pragma unroll
  for (int ii = 0; ii < 3; ii++) {
    double a0 = Pj[0] - Pi[3 * ii], a1 = Pj[1] - Pi[3 * ii + 1], a2 = Pj[2] - Pi[3 * ii + 2];
    double b0 = Pj[3] - Pi[3 * ii], b1 = Pj[4] - Pi[3 * ii + 1], b2 = Pj[5] - Pi[3 * ii + 2];
    double c0 = Pj[6] - Pi[3 * ii], c1 = Pj[7] - Pi[3 * ii + 1], c2 = Pj[8] - Pi[3 * ii + 2];
    const double la = fma(a2, a2, fma(a1, a1, a0 * a0)), lb = fma(b2, b2, fma(b1, b1, b0 * b0)), lc = fma(c2, c2, fma(c1, c1, c0 * c0));
    d2min = fmin(d2min, fmin(la, fmin(lb, lc)));
    d2max = fmax(d2max, fmax(la, fmax(lb, lc)));
    if (!hit) {
      // |v| < 1e-10 <=> |v|^2 < 1e-20: vertices of two meshes either coincide or are many orders apart
      if (la < 1.e-20) hit = true;
      else if (lb < 1.e-20) hit = true;
      else {
        const double ra = rsqrt_fast(la);
        cmin = fmin(cmin, fma(a2, b2, fma(a1, b1, a0 * b0)) * ra * rsqrt_fast(lb));
        if (lc < 1.e-20) hit = true;
        else cmin = fmin(cmin, fma(a2, c2, fma(a1, c1, a0 * c0)) * ra * rsqrt_fast(lc));
      }
    }
  }
  if (fabs(cmin - cA) < band || fabs(cmin - cB) < band) return mf_class_exact(Pi, Pj);
  if (cmin < cA) return cmin < cB ? 2 : 1;
  // distance-ratio test (also reached when a coincident vertex stopped the loop before any large angle)
  const double t95 = 0.9025 * d2max, t75 = 0.5625 * d2max, bd = 1.e-11 * d2max;
  if (fabs(d2min - t95) < bd || fabs(d2min - t75) < bd) return mf_class_exact(Pi, Pj);
  if (d2min < t75) return 2;
  if (d2min < t95) return 1;
  return hit ? 1 : 0;
}

// far pair, thread-local: sum_ii sum_jj w_ii w_jj / |x_ii - x_jj| with the row points from the staged table
// (rule `o` of the row cell) and the column points formed on the fly (thin_wall.F90:1069-1083)
__device__ __forceinline__ double far_inline(const RowSlot& R, const double* Pj, int o) {
  const int n = c_qnp[o], off = c_qoff[o];
  const double* tab = R.pts + 4 * (off - kSwOff);
  double acc = 0.0;
  for (int jj = 0; jj < n; jj++) {
    const double* b = g_qpts + 3 * (off + jj);
    const double b0 = b[0], b1 = b[1], b2 = b[2];
    const double x = xquad(b0, b1, b2, Pj[0], Pj[3], Pj[6]), y = xquad(b0, b1, b2, Pj[1], Pj[4], Pj[7]),
                 z = xquad(b0, b1, b2, Pj[2], Pj[5], Pj[8]);
    double s = 0.0;
#pragma unroll 2
    for (int ii = 0; ii < n; ii++) {
      const double2 xy = *reinterpret_cast<const double2*>(tab + 4 * ii), zw = *reinterpret_cast<const double2*>(tab + 4 * ii + 2);
      const double dx = xy.x - x, dy = xy.y - y, dz = zw.x - z;
      s = fma(zw.y, rsqrt_fast(fma(dz, dz, fma(dy, dy, dx * dx))), s);
    }
    acc = fma(tab[4 * jj + 3], s, acc);
  }
  return acc;
}

// near pair, whole warp: sum_q w_q phipot(row triangle, x_q(column cell)) with rule o (thin_wall.F90:1061-1068)
__device__ __forceinline__ double near_warp(const double* Pi, const double* nhi, const double* Pj, int o, int lane) {
  const int n = c_qnp[o], off = c_qoff[o];
  double acc = 0.0;
  for (int base = 0; base < n; base += 32) {
    const int q = base + lane;
    double v = 0.0;
    if (q < n) {
      const double* b = g_qpts + 3 * (off + q);
      const double b0 = b[0], b1 = b[1], b2 = b[2];
      const double x = xquad(b0, b1, b2, Pj[0], Pj[3], Pj[6]), y = xquad(b0, b1, b2, Pj[1], Pj[4], Pj[7]),
                   z = xquad(b0, b1, b2, Pj[2], Pj[5], Pj[8]);
      v = g_qwts[off + q] * phipot(Pi, nhi, x, y, z);
    }
    acc += warp_sum(v);
  }
  return acc;
}

__device__ __forceinline__ void stage_rows(SweepSmem& S, const SweepArgs& A, int r0, int nr, int tid, bool with_J) {
  for (int k = tid; k < nr * 16; k += kSwT) {
    const int s = k >> 4, q = k & 15;
    const int c = A.row_cells ? A.row_cells[r0 + s] : r0 + s;
    if (q < 9) S.row[s].P[q] = A.Pr[9 * (size_t)c + q];
    else if (q == 9) S.row[s].area = A.Ar[c];
    else if (q < 13) S.row[s].nrm[q - 10] = A.Nr ? A.Nr[3 * (size_t)c + q - 10] : 0.0;
  }
  if (with_J)
    for (int k = tid; k < nr * kMfQ * 3; k += kSwT) (&S.J[0][0][0])[k] = A.J[(size_t)r0 * kMfQ * 3 + k];
  __syncthreads();
  if (tid < nr) tri_normal(S.row[tid].P, S.row[tid].nh);
  for (int k = tid; k < nr * kSwPts; k += kSwT) {
    const int s = k / kSwPts, q = k - s * kSwPts;
    const double* b = g_qpts + 3 * (kSwOff + q);
    const double* P = S.row[s].P;
    S.row[s].pts[4 * q] = xquad(b[0], b[1], b[2], P[0], P[3], P[6]);
    S.row[s].pts[4 * q + 1] = xquad(b[0], b[1], b[2], P[1], P[4], P[7]);
    S.row[s].pts[4 * q + 2] = xquad(b[0], b[1], b[2], P[2], P[5], P[8]);
    S.row[s].pts[4 * q + 3] = g_qwts[kSwOff + q];
  }
  __syncthreads();
}

// MODE 0: T(i,j) with the order rule of tw_compute_Lmatblock/-Hole (= tw_compute_LmatDirect's), stored
// MODE 1: tw_compute_Lmat_MF's classes, consumed at once: F_j[q] += T(i,j) J_i[q]
template <int MODE>
__global__ void __launch_bounds__(kSwT, 2) pair_sweep_kernel(const SweepArgs A) {
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  SweepSmem& S = *reinterpret_cast<SweepSmem*>(sweep_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jx = blockIdx.x * kSwT + tid;
  const bool valid = jx < A.ncc;
  double Pj[9], area_j = 0.0;
  {
    const int c = valid ? (A.col_items ? A.col_items[jx] : jx) : 0;
#pragma unroll
    for (int q = 0; q < 9; q++) Pj[q] = valid ? A.Pc[9 * (size_t)c + q] : (double)(q + 1);
    if (valid) area_j = A.Ac[c];
  }
  double F[MODE == 1 ? kMfQ * 3 : 1];
#pragma unroll
  for (int q = 0; q < (MODE == 1 ? kMfQ * 3 : 1); q++) F[q] = 0.0;
  unsigned int nfar = 0, nclose = 0, nnear = 0;
  const int ry0 = A.row0 + blockIdx.y * A.rows_per_y, ry1 = min(A.row1, ry0 + A.rows_per_y);
  for (int r0 = ry0; r0 < ry1; r0 += kSwR) {
    const int nr = min(kSwR, ry1 - r0);
    __syncthreads();
    stage_rows(S, A, r0, nr, tid, MODE == 1);
    int nlist = 0;
    for (int s = 0; s < nr; s++) {
      const RowSlot& R = S.row[s];
      int o = 0;
      bool near = false;
      if (valid) {
        if (MODE == 0) {
          double d2min = 1.e300, d2max = 0.0;
#pragma unroll
          for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
              const double dx = R.P[3 * a] - Pj[3 * b], dy = R.P[3 * a + 1] - Pj[3 * b + 1], dz = R.P[3 * a + 2] - Pj[3 * b + 2];
              const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
              d2min = fmin(d2min, d2);
              d2max = fmax(d2max, d2);
            }
          const double floor2 = fmax(R.area, area_j) * 2.0;
          o = iquad_fast(d2min, fmax(d2max, floor2));
          if (o < 0) o = iquad_exact(R.P, Pj, 3, 3, floor2);
          near = o > 10;
        } else {
          const int cls = mf_class(R.P, Pj);
          o = cls == 0 ? 6 : 10;
          near = cls == 2;
          nclose += cls == 1;
        }
        if (!near) {
          const double T = far_inline(R, Pj, o) * R.area * area_j;
          nfar++;
          if (MODE == 0) A.T[(size_t)(r0 - A.row0 + s) * A.ldT + jx] = T;
          else {
#pragma unroll
            for (int q = 0; q < kMfQ * 3; q++) F[q] = fma(T, (&S.J[s][0][0])[q], F[q]);
          }
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, near);
      if (m == 0) continue;
      const int rank = __popc(m & ((1u << lane) - 1u));
      if (MODE == 0 && A.near_list) {  // hand the pair to the near kernel (one atomic per warp)
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(A.near_count, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (near && base + rank < A.near_cap) {
          A.near_list[base + rank] = make_int2(r0 + s, jx);
          near = false;
        }
        const unsigned m2 = __ballot_sync(0xffffffffu, near);  // what did not fit is evaluated here
        if (near) S.list[warp][nlist + __popc(m2 & ((1u << lane) - 1u))] = (uint16_t)(s << 5 | lane);
        nlist += __popc(m2);
        continue;
      }
      if (near) S.list[warp][nlist + rank] = (uint16_t)(s << 5 | lane);
      nlist += __popc(m);
    }
    __syncwarp();
    for (int e = 0; e < nlist; e++) {
      const int ent = S.list[warp][e], s = ent >> 5, owner = ent & 31;
      double Po[9];
#pragma unroll
      for (int q = 0; q < 9; q++) Po[q] = __shfl_sync(0xffffffffu, Pj[q], owner);
      int o = 10;
      if (MODE == 0) {
        // the owner's order: recomputed by the warp from the same inputs (same code path, same result)
        const RowSlot& R = S.row[s];
        const double aj = __shfl_sync(0xffffffffu, area_j, owner);
        o = iquad_exact(R.P, Po, 3, 3, fmax(R.area, aj) * 2.0);
      }
      const double T = near_warp(S.row[s].P, S.row[s].nh, Po, o, lane) * __shfl_sync(0xffffffffu, area_j, owner);
      if (lane == owner) {
        nnear++;
        if (MODE == 0) A.T[(size_t)(r0 - A.row0 + s) * A.ldT + jx] = T;
        else {
#pragma unroll
          for (int q = 0; q < kMfQ * 3; q++) F[q] = fma(T, (&S.J[s][0][0])[q], F[q]);
        }
      }
    }
  }
  if (MODE == 1 && valid) {
    double* f = A.F + ((size_t)blockIdx.y * A.ncc + jx) * (kMfQ * 3);
#pragma unroll
    for (int q = 0; q < kMfQ * 3; q++) f[q] = F[q];
  }
  if (A.counts) {
    nfar = (unsigned)warp_sum((double)(nfar - (MODE == 1 ? nclose : 0)));
    nclose = (unsigned)warp_sum((double)nclose);
    nnear = (unsigned)warp_sum((double)nnear);
    if (lane == 0) {
      atomicAdd(A.counts, (unsigned long long)nfar);
      atomicAdd(A.counts + 1, (unsigned long long)nclose);
      atomicAdd(A.counts + 2, (unsigned long long)nnear);
    }
  }
}

// near pairs of a stored-T sweep, one warp per pair over the whole device (thin_wall.F90:1061-1068)
__global__ void __launch_bounds__(256) near_pairs_kernel(const SweepArgs A) {
  const int lane = threadIdx.x & 31;
  const unsigned n = min(*A.near_count, A.near_cap);
  for (unsigned e = blockIdx.x * 8 + (threadIdx.x >> 5); e < n; e += gridDim.x * 8) {
    const int2 ent = A.near_list[e];
    const int ci = A.row_cells ? A.row_cells[ent.x] : ent.x, cj = A.col_items ? A.col_items[ent.y] : ent.y;
    double Pi[9], Pj[9], nh[3];
#pragma unroll
    for (int q = 0; q < 9; q++) {
      Pi[q] = A.Pr[9 * (size_t)ci + q];
      Pj[q] = A.Pc[9 * (size_t)cj + q];
    }
    tri_normal(Pi, nh);
    const double ai = A.Ar[ci], aj = A.Ac[cj];
    const int o = iquad_exact(Pi, Pj, 3, 3, fmax(ai, aj) * 2.0);
    const double T = near_warp(Pi, nh, Pj, o, lane) * aj;
    if (lane == 0) A.T[(size_t)(ent.x - A.row0) * A.ldT + ent.y] = T;
  }
}

// MODE 2 of the sweep: B operator of a row cell at a mesh vertex (thin_wall_hodlr.F90:612-676 = thin_wall.F90:2030-2088).
// D(i,p) such that the contribution of vertex k of cell i is D x qbasis(:,k,i):
//   far : D = -area_i sum_q w_q d_q/|d_q|^3, d_q = r_p - x_q      near: D = grad phi by central differences
__global__ void __launch_bounds__(kSwT) bops_sweep_kernel(const SweepArgs A) {
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  SweepSmem& S = *reinterpret_cast<SweepSmem*>(sweep_smem);
  const int tid = threadIdx.x;
  const int jx = blockIdx.x * kSwT + tid;
  const bool valid = jx < A.ncc;
  double X[3] = {0.0, 0.0, 0.0}, vaj = 0.0;
  if (valid) {
    const int p = A.col_items ? A.col_items[jx] : jx;
    X[0] = A.rc[3 * (size_t)p];
    X[1] = A.rc[3 * (size_t)p + 1];
    X[2] = A.rc[3 * (size_t)p + 2];
    vaj = A.vac[p];
  }
  const double pi = 3.14159265358979323846;
  const int ry0 = A.row0 + blockIdx.y * A.rows_per_y, ry1 = min(A.row1, ry0 + A.rows_per_y);
  for (int r0 = ry0; r0 < ry1; r0 += kSwR) {
    const int nr = min(kSwR, ry1 - r0);
    __syncthreads();
    if (tid == 0) S.bcount = 0;
    stage_rows(S, A, r0, nr, tid, false);
    if (valid)
      for (int s = 0; s < nr; s++) {
        const RowSlot& R = S.row[s];
        const double floor2 = fmax(R.area, __ddiv_rn(vaj, __dmul_rn(pi, pi)));
        double d2min = 1.e300, d2max = 0.0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const double dx = R.P[3 * a] - X[0], dy = R.P[3 * a + 1] - X[1], dz = R.P[3 * a + 2] - X[2];
          const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
          d2min = fmin(d2min, d2);
          d2max = fmax(d2max, d2);
        }
        int o = iquad_fast(d2min, fmax(d2max, floor2));
        if (o < 0) o = iquad_exact(R.P, X, 3, 1, floor2);
        if (o > 10) {
          // on-surface vertex (dl_min < 1e-8): the exact test decides, like the order
          const bool nb = d2min < 1.0000001e-16 && (d2min < 0.9999999e-16 || iquad_exact(R.P, X, 3, 1, floor2) == 18);
          S.blist[atomicAdd(&S.bcount, 1)] = (unsigned)s << 16 | (nb ? 1u << 15 : 0u) | (unsigned)tid;
          continue;
        }
        const int n = c_qnp[o], off = c_qoff[o];
        const double* tab = R.pts + 4 * (off - kSwOff);
        double h0 = 0.0, h1 = 0.0, h2 = 0.0;
        for (int q = 0; q < n; q++) {
          const double dx = X[0] - tab[4 * q], dy = X[1] - tab[4 * q + 1], dz = X[2] - tab[4 * q + 2];
          const double ri = rsqrt_fast(fma(dz, dz, fma(dy, dy, dx * dx)));
          const double w3 = tab[4 * q + 3] * (ri * ri * ri);
          h0 = fma(w3, dx, h0);
          h1 = fma(w3, dy, h1);
          h2 = fma(w3, dz, h2);
        }
        double* D = A.T + ((size_t)(r0 - A.row0 + s) * A.ldT + jx) * 3;
        D[0] = -h0 * R.area;
        D[1] = -h1 * R.area;
        D[2] = -h2 * R.area;
      }
    __syncthreads();
    for (int k = tid; k < S.bcount; k += kSwT) {
      const unsigned e = S.blist[k];
      const int s = e >> 16, t = e & 0x7fff;
      const bool nb = (e >> 15) & 1u;
      const int jy = blockIdx.x * kSwT + t;
      const int p = A.col_items ? A.col_items[jy] : jy;
      const RowSlot& R = S.row[s];
      const double B_dx = 1.e-6;
      double pt[3] = {A.rc[3 * (size_t)p], A.rc[3 * (size_t)p + 1], A.rc[3 * (size_t)p + 2]}, diff[3] = {0.0, 0.0, 0.0};
      if (nb)
        for (int d = 0; d < 3; d++) pt[d] = xsub(pt[d], xmul(xmul(R.nrm[d], 10.0), B_dx));
      for (int ik = 1; ik <= 2; ik++) {
        if (ik == 2)
          for (int d = 0; d < 3; d++) pt[d] = xadd(pt[d], xmul(xmul(R.nrm[d], 20.0), B_dx));
#pragma unroll
        for (int jj = 0; jj < 3; jj++) {
          pt[jj] = xadd(pt[jj], B_dx);
          double tmp = phipot(R.P, R.nh, pt[0], pt[1], pt[2]);
          diff[jj] = xadd(diff[jj], __ddiv_rn(tmp, 2.0 * B_dx));
          pt[jj] = xsub(pt[jj], 2.0 * B_dx);
          tmp = phipot(R.P, R.nh, pt[0], pt[1], pt[2]);
          diff[jj] = xsub(diff[jj], __ddiv_rn(tmp, 2.0 * B_dx));
          pt[jj] = xadd(pt[jj], B_dx);
        }
        if (!nb) break;
      }
      if (nb)
        for (int d = 0; d < 3; d++) diff[d] = diff[d] / 2.0;
      double* D = A.T + ((size_t)(r0 - A.row0 + s) * A.ldT + jy) * 3;
      D[0] = diff[0];
      D[1] = diff[1];
      D[2] = diff[2];
    }
  }
}

// ---- contraction of a stored T block: out[a][b] (+)= 1/(4 pi) sum_{(r,k) in inc(a)} sum_{(c,l) in inc(b)}
// s_a s_b (qbasis_r[k] . qbasis_c[l]) T[r][c], incidences in ascending list order (fixed summation order).
// Incidence word: local list index << 3 | negative << 2 | local vertex.
struct ContractArgs {
  const int *kri, *lri, *kci, *lci;
  const double *Er, *Ec;              // qbasis of the row / column model [nc][3][3]
  const int *row_cells, *col_cells;   // list index -> cell id (NULL: identity)
  const double* T;
  long long ldT;
  int r0, r1;                         // row-list range held by T
  int nrd, ncd;
  double* out;
  long long ld;
  int accumulate;
};
__global__ void block_contract_kernel(const ContractArgs A) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)A.nrd * A.ncd) return;
  const int a = (int)(idx / A.ncd), b = (int)(idx - (long long)a * A.ncd);
  double sum = 0.0;
  for (int i1 = A.kri[a]; i1 < A.kri[a + 1]; i1++) {
    const int w1 = A.lri[i1], r = w1 >> 3;
    if (r < A.r0 || r >= A.r1) continue;
    const double* e1 = A.Er + 9 * (size_t)(A.row_cells ? A.row_cells[r] : r) + 3 * (w1 & 3);
    const double* Trow = A.T + (size_t)(r - A.r0) * A.ldT;
    for (int i2 = A.kci[b]; i2 < A.kci[b + 1]; i2++) {
      const int w2 = A.lci[i2], c = w2 >> 3;
      const double* e2 = A.Ec + 9 * (size_t)(A.col_cells ? A.col_cells[c] : c) + 3 * (w2 & 3);
      double v = (e1[0] * e2[0] + e1[1] * e2[1] + e1[2] * e2[2]) * Trow[c];
      if (((w1 ^ w2) & 4) != 0) v = -v;
      sum += v;
    }
  }
  const double v = sum / (4.0 * 3.14159265358979323846);
  double* o = A.out + (size_t)a * A.ld + b;
  *o = A.accumulate ? *o + v : v;
}

// B operator: out[a][p] (+)= 1/(4 pi) sum_{(r,k) in inc(a)} (D[r][p] x qbasis_r[k])_dir   (dir < 0: all three, out[3][nrd][ld])
struct BContractArgs {
  const int *kri, *lri;
  const double* Er;
  const int* row_cells;
  const double* D;
  long long ldT;
  int r0, r1, nrd, ncp, dir;
  double* out;
  long long ld, comp_stride;
  int accumulate;
};
__global__ void bops_contract_kernel(const BContractArgs A) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)A.nrd * A.ncp) return;
  const int a = (int)(idx / A.ncp), p = (int)(idx - (long long)a * A.ncp);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int i1 = A.kri[a]; i1 < A.kri[a + 1]; i1++) {
    const int w1 = A.lri[i1], r = w1 >> 3;
    if (r < A.r0 || r >= A.r1) continue;
    const double* e = A.Er + 9 * (size_t)(A.row_cells ? A.row_cells[r] : r) + 3 * (w1 & 3);
    const double* d = A.D + ((size_t)(r - A.r0) * A.ldT + p) * 3;
    double c0 = d[1] * e[2] - d[2] * e[1], c1 = d[2] * e[0] - d[0] * e[2], c2 = d[0] * e[1] - d[1] * e[0];
    if (w1 & 4) { c0 = -c0; c1 = -c1; c2 = -c2; }
    s0 += c0;
    s1 += c1;
    s2 += c2;
  }
  const double sc = 4.0 * 3.14159265358979323846;
  for (int k = 0; k < 3; k++) {
    if (A.dir >= 0 && A.dir != k) continue;
    const double v = (k == 0 ? s0 : k == 1 ? s1 : s2) / sc;
    double* o = A.out + (A.dir >= 0 ? 0 : (size_t)k * A.comp_stride) + (size_t)a * A.ld + p;
    *o = A.accumulate ? *o + v : v;
  }
}

// ---- matrix-free apply: row currents and the gather of the column DOFs ---------------------------------------
// J[i][q] = sum_k qbasis(:,k,i) a(pmap(lc(k,i)),q) + sum_holes sign qbasis(:,lv,i) a(np_active + hole, q)
__global__ void mf_rowcur_kernel(int nc, const int* __restrict__ lc, const int* __restrict__ pmap, const int* __restrict__ kfh,
                                 const int* __restrict__ lfh, int np_active, const double* __restrict__ E,
                                 const double* __restrict__ a, long long nelems, int q0, int nq, double* __restrict__ J) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  for (int q = 0; q < kMfQ; q++) {
    double j0 = 0.0, j1 = 0.0, j2 = 0.0;
    if (q < nq) {
      const double* aq = a + (size_t)(q0 + q) * nelems;
      for (int k = 0; k < 3; k++) {
        const int ik = pmap[lc[3 * i + k]];
        if (ik == 0) continue;
        const double v = aq[ik - 1];
        j0 = fma(E[9 * (size_t)i + 3 * k], v, j0);
        j1 = fma(E[9 * (size_t)i + 3 * k + 1], v, j1);
        j2 = fma(E[9 * (size_t)i + 3 * k + 2], v, j2);
      }
      for (int ii = kfh[i]; ii < kfh[i + 1]; ii++) {
        const int h = lfh[2 * ii], k = lfh[2 * ii + 1];
        const double v = (h < 0 ? -1.0 : 1.0) * aq[np_active + abs(h) - 1];
        j0 = fma(E[9 * (size_t)i + 3 * k], v, j0);
        j1 = fma(E[9 * (size_t)i + 3 * k + 1], v, j1);
        j2 = fma(E[9 * (size_t)i + 3 * k + 2], v, j2);
      }
    }
    J[((size_t)i * kMfQ + q) * 3] = j0;
    J[((size_t)i * kMfQ + q) * 3 + 1] = j1;
    J[((size_t)i * kMfQ + q) * 3 + 2] = j2;
  }
}
// b(jk,q) = 1/(4 pi) sum_{(j,l) in inc(jk)} s qbasis(:,l,j) . (sum_y F[y][j][q])
__global__ void mf_gather_kernel(int ndof, const int* __restrict__ kdi, const int* __restrict__ ldi, const double* __restrict__ E,
                                 const double* __restrict__ F, int ny, int ncc, long long nelems, int q0, int nq,
                                 double* __restrict__ b) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ndof * nq) return;
  const int q = t / ndof, jk = t - q * ndof;
  double sum = 0.0;
  for (int i1 = kdi[jk]; i1 < kdi[jk + 1]; i1++) {
    const int w = ldi[i1], j = w >> 3;
    const double* e = E + 9 * (size_t)j + 3 * (w & 3);
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    for (int y = 0; y < ny; y++) {
      const double* f = F + (((size_t)y * ncc + j) * kMfQ + q) * 3;
      f0 += f[0];
      f1 += f[1];
      f2 += f[2];
    }
    double v = e[0] * f0 + e[1] * f1 + e[2] * f2;
    if (w & 4) v = -v;
    sum += v;
  }
  b[(size_t)(q0 + q) * nelems + jk] = sum / (4.0 * 3.14159265358979323846);
}

// ---- reduced model: Y[q][r] = sum_j A[r][j] X[q][j] for up to 8 vectors per pass over the rows (HBM-bound: the row
// block is read once per 8 vectors), and the small Gram products G[a][b] = sum_i U[a][i] W[b][i] --------------------
constexpr int kRedQ = 8;
__global__ void __launch_bounds__(256) rows_apply_multi_kernel(const double* __restrict__ A, long long ld, int nrows, int n,
                                                               const double* __restrict__ X, long long ldx, int nq,
                                                               double* __restrict__ Y, long long ldy) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= nrows) return;
  const double* a = A + (long long)row * ld;
  double s[kRedQ];
#pragma unroll
  for (int q = 0; q < kRedQ; q++) s[q] = 0.0;
  for (int j = lane; j < n; j += 32) {
    const double v = __ldcs(a + j);
#pragma unroll
    for (int q = 0; q < kRedQ; q++)
      if (q < nq) s[q] = fma(v, __ldg(X + (size_t)q * ldx + j), s[q]);
  }
#pragma unroll
  for (int q = 0; q < kRedQ; q++) {
    const double t = warp_sum(s[q]);
    if (lane == 0 && q < nq) Y[(size_t)q * ldy + row] = t;
  }
}
__global__ void __launch_bounds__(256) gram_kernel(const double* __restrict__ U, long long ldu, const double* __restrict__ W,
                                                   long long ldw, int n, int nb, double* __restrict__ G) {
  // G[a][b], a = blockIdx.x / nb
  __shared__ double part[8];
  const int a = blockIdx.x / nb, b = blockIdx.x - a * nb;
  const double *u = U + (size_t)a * ldu, *w = W + (size_t)b * ldw;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s = fma(u[i], w[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; k++) t += part[k];
    G[blockIdx.x] = t;
  }
}

}  // namespace twk

// =================================================================================================================
namespace tw {

// Device mirror of the plain per-cell arrays of a model for the list sweeps (kept between calls: ACA asks for
// thousands of strips of the same model).
struct BlockCtx {
  int device = -1;
  DBuf<double> P, A, E, N, r, va;
  DBuf<int> lc, pmap, kfh, lfh, kdi, ldi;
  int ndof = 0;
  std::string up(const Model& m) {
    std::string e;
    CellArrays ca;
    if (!(e = ca.up(m)).empty()) return e;
    std::swap(P.p, ca.P.p);
    std::swap(A.p, ca.A.p);
    std::swap(E.p, ca.E.p);
    if (!(e = N.up(m.norm)).empty()) return e;
    if (!(e = r.up(m.r)).empty()) return e;
    if (!(e = va.up(m.va)).empty()) return e;
    if (!(e = lc.up(m.lc)).empty()) return e;
    if (!(e = pmap.up(m.pmap)).empty()) return e;
    if (!(e = kfh.up(m.kfh)).empty()) return e;
    std::vector<int> l = m.lfh;
    if (l.empty()) l.assign(2, 0);
    if (!(e = lfh.up(l)).empty()) return e;
    std::vector<int> kd, ld_;
    dof_incidence(m, kd, ld_);
    ndof = (int)kd.size() - 1;
    if (!(e = kdi.up(kd)).empty()) return e;
    return ldi.up(ld_);
  }
};

static std::string block_ctx(Model& m, BlockCtx*& out) {
  std::string e = need_gpu();
  if (!e.empty()) return e;
  int dev = 0;
  CKO(cudaGetDevice(&dev));
  auto ctx = std::static_pointer_cast<BlockCtx>(m.block_ctx);
  if (!ctx || ctx->device != dev) {
    ctx = std::make_shared<BlockCtx>();
    if (!(e = ctx->up(m)).empty()) return e;
    ctx->device = dev;
    m.block_ctx = ctx;
  }
  out = ctx.get();
  static thread_local int attr_dev = -1;
  if (attr_dev != dev) {
    CKO(cudaFuncSetAttribute(twk::pair_sweep_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(twk::SweepSmem)));
    CKO(cudaFuncSetAttribute(twk::pair_sweep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(twk::SweepSmem)));
    CKO(cudaFuncSetAttribute(twk::bops_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(twk::SweepSmem)));
    attr_dev = dev;
  }
  return "";
}

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// vertex subset -> cells touching it (ascending) + incidence CSR of the block's points over that cell list
// (oft_tw_block: icell / inv_map, thin_wall_hodlr.F90:843-870)
static std::string vertex_block(const Model& m, int npts, const int* pts, std::vector<int>& cells, std::vector<int>& kinc,
                                std::vector<int>& linc) {
  std::vector<int> inv(m.np, -1);
  for (int k = 0; k < npts; k++) {
    if (pts[k] < 0 || pts[k] >= m.np) return "vertex id out of range";
    if (inv[pts[k]] >= 0) return "duplicate vertex id";
    inv[pts[k]] = k;
  }
  cells.clear();
  for (int k = 0; k < npts; k++)
    for (int i = m.kpc[pts[k]]; i < m.kpc[pts[k] + 1]; i++) cells.push_back(m.lpc[i]);
  std::sort(cells.begin(), cells.end());
  cells.erase(std::unique(cells.begin(), cells.end()), cells.end());
  kinc.assign(npts + 1, 0);
  for (size_t ci = 0; ci < cells.size(); ci++)
    for (int k = 0; k < 3; k++) {
      const int b = inv[m.lc[3 * cells[ci] + k]];
      if (b >= 0) kinc[b + 1]++;
    }
  for (int k = 0; k < npts; k++) kinc[k + 1] += kinc[k];
  linc.resize(kinc[npts]);
  std::vector<int> fill(kinc.begin(), kinc.end() - 1);
  for (size_t ci = 0; ci < cells.size(); ci++)
    for (int k = 0; k < 3; k++) {
      const int b = inv[m.lc[3 * cells[ci] + k]];
      if (b >= 0) linc[fill[b]++] = ((int)ci << 3) | k;
    }
  return "";
}

static void sweep_grid(int ncc, int nrows, dim3& grid, int& rows_per_y) {
  const int nx = (ncc + twk::kSwT - 1) / twk::kSwT;
  const int groups = (nrows + twk::kSwR - 1) / twk::kSwR;
  int ny = std::max(1, std::min(groups, (4 * 148 + nx - 1) / nx));
  ny = std::min(ny, 65535);
  rows_per_y = ((groups + ny - 1) / ny) * twk::kSwR;
  ny = (nrows + rows_per_y - 1) / rows_per_y;
  grid = dim3(nx, ny);
}

// Common driver of the stored-T builds: T in row-list chunks of <= 256 MB, contracted chunk by chunk in list order.
static std::string stored_sweep(BlockCtx& R, BlockCtx& C, const DBuf<int>* row_cells, int nrc, const DBuf<int>* col_cells, int ncc,
                                const DBuf<int>& kri, const DBuf<int>& lri, int nrd, const int* kci, const int* lci, int ncd,
                                double* d_out, long long ld, cudaStream_t stream) {
  if (nrd == 0 || ncd == 0) return "";
  const long long ldT = ((long long)ncc + 3) & ~3ll;
  const int chunk = (int)std::max<long long>(twk::kSwR, std::min<long long>(nrc, ((256ll << 20) / 8) / std::max<long long>(ldT, 1)));
  double* T = nullptr;
  CKO(cudaMallocAsync((void**)&T, (size_t)chunk * ldT * 8, stream));
  // near pairs of a chunk: listed by the sweep, evaluated by a second kernel over the whole device
  const unsigned near_cap = 1u << 21;
  int2* near_list = nullptr;
  unsigned* near_count = nullptr;
  CKO(cudaMallocAsync((void**)&near_list, (size_t)near_cap * sizeof(int2) + sizeof(unsigned), stream));
  near_count = reinterpret_cast<unsigned*>(near_list + near_cap);
  int nsm = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  for (int r0 = 0, it = 0; r0 < std::max(nrc, 1); r0 += chunk, it++) {
    const int r1 = std::min(nrc, r0 + chunk);
    if (r1 > r0) {
      twk::SweepArgs a{};
      a.near_list = near_list;
      a.near_count = near_count;
      a.near_cap = near_cap;
      CKO(cudaMemsetAsync(near_count, 0, sizeof(unsigned), stream));
      a.Pr = R.P.p; a.Ar = R.A.p; a.Nr = R.N.p;
      a.Pc = C.P.p; a.Ac = C.A.p;
      a.row_cells = row_cells ? row_cells->p : nullptr;
      a.col_items = col_cells ? col_cells->p : nullptr;
      a.nrc = nrc; a.ncc = ncc; a.row0 = r0; a.row1 = r1;
      a.T = T; a.ldT = ldT;
      dim3 grid;
      sweep_grid(ncc, r1 - r0, grid, a.rows_per_y);
      twk::pair_sweep_kernel<0><<<grid, twk::kSwT, sizeof(twk::SweepSmem), stream>>>(a);
      CKO(cudaGetLastError());
      note_launch();
      twk::near_pairs_kernel<<<4 * nsm, 256, 0, stream>>>(a);
      CKO(cudaGetLastError());
      note_launch();
    }
    twk::ContractArgs c{};
    c.kri = kri.p; c.lri = lri.p; c.kci = kci; c.lci = lci;
    c.Er = R.E.p; c.Ec = C.E.p;
    c.row_cells = row_cells ? row_cells->p : nullptr;
    c.col_cells = col_cells ? col_cells->p : nullptr;
    c.T = T; c.ldT = ldT; c.r0 = r0; c.r1 = r1; c.nrd = nrd; c.ncd = ncd;
    c.out = d_out; c.ld = ld; c.accumulate = it > 0;
    const long long nt = (long long)nrd * ncd;
    twk::block_contract_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, stream>>>(c);
    CKO(cudaGetLastError());
    note_launch();
  }
  CKO(cudaFreeAsync(T, stream));
  CKO(cudaFreeAsync(near_list, stream));
  return "";
}

// out may be host or device memory; returns a device pointer to write into (staging buffer if host)
struct OutStage {
  double* d = nullptr;
  double* host = nullptr;
  size_t bytes = 0;
  bool staged = false;
  std::string begin(double* out, size_t count, cudaStream_t st) {
    bytes = count * 8;
    if (is_device_ptr(out)) {
      d = out;
      return "";
    }
    host = out;
    staged = true;
    CKO(cudaMallocAsync((void**)&d, std::max<size_t>(bytes, 8), st));
    return "";
  }
  std::string end(cudaStream_t st) {
    if (!staged) return "";
    CKO(cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, st));
    CKO(cudaStreamSynchronize(st));
    CKO(cudaFreeAsync(d, st));
    return "";
  }
};

// tw_compute_Lmatblock (thin_wall_hodlr.F90:289-404): out[a][b] = Lmat(col_pts[b], row_pts[a]) for two vertex blocks
std::string gpu_lmatblock(Model& mr, Model& mc, int nrp, const int* row_pts, int ncp, const int* col_pts, double* out, long long ld,
                          cudaStream_t stream) {
  if (ld < ncp) return "thincurr_b200_Lmatblock: ld < number of column points";
  BlockCtx *R = nullptr, *C = nullptr;
  std::string e;
  if (!(e = block_ctx(mr, R)).empty()) return e;
  if (!(e = block_ctx(mc, C)).empty()) return e;
  std::vector<int> rc, kr, lr, cc, kc, lcn;
  if (!(e = vertex_block(mr, nrp, row_pts, rc, kr, lr)).empty()) return "thincurr_b200_Lmatblock: row block: " + e;
  if (!(e = vertex_block(mc, ncp, col_pts, cc, kc, lcn)).empty()) return "thincurr_b200_Lmatblock: column block: " + e;
  DBuf<int> d_rc, d_kr, d_lr, d_cc, d_kc, d_lc;
  if (!(e = d_rc.up(rc)).empty() || !(e = d_kr.up(kr)).empty() || !(e = d_lr.up(lr)).empty() || !(e = d_cc.up(cc)).empty() ||
      !(e = d_kc.up(kc)).empty() || !(e = d_lc.up(lcn)).empty())
    return e;
  OutStage os;
  if (!(e = os.begin(out, (size_t)nrp * ld, stream)).empty()) return e;
  if (os.staged) CKO(cudaMemsetAsync(os.d, 0, os.bytes, stream));
  if (!(e = stored_sweep(*R, *C, &d_rc, (int)rc.size(), &d_cc, (int)cc.size(), d_kr, d_lr, nrp, d_kc.p, d_lc.p, ncp, os.d, ld, stream))
           .empty())
    return e;
  if (!(e = os.end(stream)).empty()) return e;
  CKO(cudaStreamSynchronize(stream));  // the index lists above are freed on return
  return "";
}

// tw_compute_LmatHole(self,self) (thin_wall_hodlr.F90:136-285): out[h][:] = Lmat(:, h) for the hole and V-coil columns
std::string gpu_lmathole(Model& m, double* out, long long ld, cudaStream_t stream) {
  if (ld < m.nelems) return "thincurr_b200_LmatHole: ld < nelems";
  if (m.n_vcoils > 0 && !m.have_coil_mutuals) return "Coil mutuals required if, # of Vcoils > 0";
  const int nk = m.nholes + m.n_vcoils;
  if (nk == 0) return "";
  BlockCtx* R = nullptr;
  std::string e;
  if (!(e = block_ctx(m, R)).empty()) return e;
  // row list: the cells with hole entries; row DOFs: the holes, with (list index, sign, local vertex) incidences
  std::vector<int> rc, kr(m.nholes + 1, 0), lr;
  std::vector<int> idx(m.nc, -1);
  for (int c = 0; c < m.nc; c++)
    if (m.kfh[c + 1] > m.kfh[c]) {
      idx[c] = (int)rc.size();
      rc.push_back(c);
    }
  for (int c : rc)
    for (int ii = m.kfh[c]; ii < m.kfh[c + 1]; ii++) kr[std::abs(m.lfh[2 * ii])]++;
  for (int h = 0; h < m.nholes; h++) kr[h + 1] += kr[h];
  lr.resize(kr[m.nholes]);
  {
    std::vector<int> fill(kr.begin(), kr.end() - 1);
    for (int c : rc)
      for (int ii = m.kfh[c]; ii < m.kfh[c + 1]; ii++) {
        const int h = m.lfh[2 * ii];
        lr[fill[std::abs(h) - 1]++] = (idx[c] << 3) | (h < 0 ? 4 : 0) | m.lfh[2 * ii + 1];
      }
  }
  DBuf<int> d_rc, d_kr, d_lr;
  if (!(e = d_rc.up(rc)).empty() || !(e = d_kr.up(kr)).empty() || !(e = d_lr.up(lr)).empty()) return e;
  OutStage os;
  if (!(e = os.begin(out, (size_t)nk * ld, stream)).empty()) return e;
  CKO(cudaMemsetAsync(os.d, 0, (size_t)nk * ld * 8, stream));
  const int ncd = m.np_active + m.nholes;
  if (!(e = stored_sweep(*R, *R, &d_rc, (int)rc.size(), nullptr, m.nc, d_kr, d_lr, m.nholes, R->kdi.p, R->ldi.p, ncd, os.d, ld, stream))
           .empty())
    return e;
  if (m.n_vcoils > 0) {  // :253-276, from Ael2coil / Acoil2coil (host, Fortran (nelems,n_vcoils) / (n_vcoils,n_vcoils))
    CKO(cudaStreamSynchronize(stream));
    const double s = 4.0 * kPi;
    const size_t ne = (size_t)m.nelems;
    std::vector<double> col(ncd);
    for (int j = 0; j < m.n_vcoils; j++) {
      for (int i = 0; i < ncd; i++) col[i] = m.Ael2coil.p[(size_t)j * ne + i] / s;
      CKO(cudaMemcpy(os.d + (size_t)(m.nholes + j) * ld, col.data(), (size_t)ncd * 8, cudaMemcpyHostToDevice));
      for (int i = 0; i < m.nholes; i++) {
        const double v = m.Ael2coil.p[(size_t)j * ne + m.np_active + i] / s;
        CKO(cudaMemcpy(os.d + (size_t)i * ld + ncd + j, &v, 8, cudaMemcpyHostToDevice));
      }
      for (int i = 0; i < m.n_vcoils; i++) {
        const double v = m.Acoil2coil.p[(size_t)j * m.n_vcoils + i] / s;
        CKO(cudaMemcpy(os.d + (size_t)(m.nholes + i) * ld + ncd + j, &v, 8, cudaMemcpyHostToDevice));
      }
    }
  }
  if (!(e = os.end(stream)).empty()) return e;
  CKO(cudaStreamSynchronize(stream));
  return "";
}

// tw_compute_Bops_block (thin_wall_hodlr.F90:580-691): out[a][b] = Bop(col_pts[b], row_pts[a]) of component dir
// (dir < 0: all three components, out[3][nrp][ld])
std::string gpu_bops_block(Model& m, int nrp, const int* row_pts, int ncp, const int* col_pts, int dir, double* out, long long ld,
                           cudaStream_t stream) {
  if (ld < ncp) return "thincurr_b200_Bops_block: ld < number of column points";
  if (dir > 2) return "thincurr_b200_Bops_block: dir must be 0, 1, 2 or negative (all)";
  for (int k = 0; k < ncp; k++)
    if (col_pts[k] < 0 || col_pts[k] >= m.np) return "thincurr_b200_Bops_block: column vertex id out of range";
  BlockCtx* R = nullptr;
  std::string e;
  if (!(e = block_ctx(m, R)).empty()) return e;
  std::vector<int> rc, kr, lr;
  if (!(e = vertex_block(m, nrp, row_pts, rc, kr, lr)).empty()) return "thincurr_b200_Bops_block: row block: " + e;
  if (nrp == 0 || ncp == 0) return "";
  DBuf<int> d_rc, d_kr, d_lr, d_cp;
  std::vector<int> cp(col_pts, col_pts + ncp);
  if (!(e = d_rc.up(rc)).empty() || !(e = d_kr.up(kr)).empty() || !(e = d_lr.up(lr)).empty() || !(e = d_cp.up(cp)).empty()) return e;
  const int ncomp = dir < 0 ? 3 : 1;
  OutStage os;
  if (!(e = os.begin(out, (size_t)ncomp * nrp * ld, stream)).empty()) return e;
  if (os.staged) CKO(cudaMemsetAsync(os.d, 0, os.bytes, stream));
  const int nrc = (int)rc.size();
  const long long ldT = ncp;
  const int chunk = (int)std::max<long long>(twk::kSwR, std::min<long long>(nrc, ((256ll << 20) / 24) / std::max<long long>(ldT, 1)));
  double* D = nullptr;
  CKO(cudaMallocAsync((void**)&D, (size_t)chunk * ldT * 24, stream));
  for (int r0 = 0, it = 0; r0 < std::max(nrc, 1); r0 += chunk, it++) {
    const int r1 = std::min(nrc, r0 + chunk);
    if (r1 > r0) {
      twk::SweepArgs a{};
      a.Pr = R->P.p; a.Ar = R->A.p; a.Nr = R->N.p;
      a.rc = R->r.p; a.vac = R->va.p;
      a.row_cells = d_rc.p; a.col_items = d_cp.p;
      a.nrc = nrc; a.ncc = ncp; a.row0 = r0; a.row1 = r1;
      a.T = D; a.ldT = ldT;
      dim3 grid;
      sweep_grid(ncp, r1 - r0, grid, a.rows_per_y);
      twk::bops_sweep_kernel<<<grid, twk::kSwT, sizeof(twk::SweepSmem), stream>>>(a);
      CKO(cudaGetLastError());
      note_launch();
    }
    twk::BContractArgs c{};
    c.kri = d_kr.p; c.lri = d_lr.p; c.Er = R->E.p; c.row_cells = d_rc.p;
    c.D = D; c.ldT = ldT; c.r0 = r0; c.r1 = r1; c.nrd = nrp; c.ncp = ncp; c.dir = dir;
    c.out = os.d; c.ld = ld; c.comp_stride = (long long)nrp * ld; c.accumulate = it > 0;
    const long long nt = (long long)nrp * ncp;
    twk::bops_contract_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, stream>>>(c);
    CKO(cudaGetLastError());
    note_launch();
  }
  CKO(cudaFreeAsync(D, stream));
  if (!(e = os.end(stream)).empty()) return e;
  CKO(cudaStreamSynchronize(stream));
  return "";
}

// tw_compute_Lmat_MF (thin_wall.F90:1190-1414): vec2[q][:] = M vec1[q][:], vec1 [nrhs][m1.nelems], vec2 [nrhs][m2.nelems]
// (HOST arrays; the V-coil parts are not computed, as in the reference :1385).  counts[3] (optional): pairs per class.
std::string gpu_cross_eval(Model& m1, Model& m2, int nrhs, const double* vec1, double* vec2, long long* counts) {
  if (nrhs <= 0) return "";
  BlockCtx *R = nullptr, *C = nullptr;
  std::string e;
  if (!(e = block_ctx(m1, R)).empty()) return e;
  if (!(e = block_ctx(m2, C)).empty()) return e;
  if ((m1.n_vcoils > 0 || m2.n_vcoils > 0) && m1.verbose) printf("WARNING: V-coil contributions were not computed.\n");
  cudaStream_t stream = nullptr;
  DBuf<double> a, b, J, F;
  DBuf<unsigned long long> cnt;
  std::vector<double> av(vec1, vec1 + (size_t)nrhs * m1.nelems);
  if (!(e = a.up(av)).empty()) return e;
  if (!(e = b.zeros((size_t)nrhs * m2.nelems)).empty()) return e;
  if (!(e = cnt.zeros(3)).empty()) return e;
  if (!(e = J.zeros((size_t)m1.nc * twk::kMfQ * 3)).empty()) return e;
  twk::SweepArgs s{};
  s.Pr = R->P.p; s.Ar = R->A.p; s.Nr = nullptr;
  s.Pc = C->P.p; s.Ac = C->A.p;
  s.nrc = m1.nc; s.ncc = m2.nc; s.row0 = 0; s.row1 = m1.nc;
  dim3 grid;
  sweep_grid(m2.nc, m1.nc, grid, s.rows_per_y);
  if (!(e = F.zeros((size_t)grid.y * m2.nc * twk::kMfQ * 3)).empty()) return e;
  s.J = J.p; s.F = F.p;
  const int ndof2 = m2.np_active + m2.nholes;
  for (int q0 = 0; q0 < nrhs; q0 += twk::kMfQ) {
    const int nq = std::min(twk::kMfQ, nrhs - q0);
    twk::mf_rowcur_kernel<<<(m1.nc + 127) / 128, 128, 0, stream>>>(m1.nc, R->lc.p, R->pmap.p, R->kfh.p, R->lfh.p, m1.np_active, R->E.p,
                                                                    a.p, m1.nelems, q0, nq, J.p);
    CKO(cudaGetLastError());
    note_launch();
    s.counts = (counts && q0 == 0) ? cnt.p : nullptr;
    twk::pair_sweep_kernel<1><<<grid, twk::kSwT, sizeof(twk::SweepSmem), stream>>>(s);
    CKO(cudaGetLastError());
    note_launch();
    if (ndof2 > 0) {
      twk::mf_gather_kernel<<<(ndof2 * nq + 255) / 256, 256, 0, stream>>>(ndof2, C->kdi.p, C->ldi.p, C->E.p, F.p, (int)grid.y, m2.nc,
                                                                          m2.nelems, q0, nq, b.p);
      CKO(cudaGetLastError());
      note_launch();
    }
  }
  CKO(cudaMemcpy(vec2, b.p, (size_t)nrhs * m2.nelems * 8, cudaMemcpyDeviceToHost));
  if (counts) {
    unsigned long long h[3];
    CKO(cudaMemcpy(h, cnt.p, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) counts[k] = (long long)h[k];
  }
  return "";
}

// ---- reduced model (tw_reduce_model, thin_wall_solvers.F90:1180-1359) ----------------------------------------------
// Y[q][:] = A X[q][:] for a HOST row-major matrix A[nrows][n] streamed once through two device slabs (all q at once
// in groups of 8), then G = U Y^T style Gram products on the device.
std::string gpu_host_matrix_multi(const double* A, size_t nrows, size_t n, int nq, const double* d_X, long long ldx, double* d_Y,
                                  long long ldy) {
  const size_t slab = std::max<size_t>(1, std::min(nrows, ((size_t)256 << 20) / (n * 8)));
  double* d_a[2] = {nullptr, nullptr};
  cudaStream_t sc = nullptr, sk = nullptr;
  cudaEvent_t up[2] = {nullptr, nullptr}, used[2] = {nullptr, nullptr};
  auto run = [&]() -> std::string {
    CKO(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
    CKO(cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CKO(cudaMalloc((void**)&d_a[i], slab * n * 8));
      CKO(cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming));
      CKO(cudaEventCreateWithFlags(&used[i], cudaEventDisableTiming));
    }
    int k = 0;
    for (size_t r0 = 0; r0 < nrows; r0 += slab, k ^= 1) {
      const size_t nr = std::min(slab, nrows - r0);
      CKO(cudaStreamWaitEvent(sc, used[k], 0));
      CKO(cudaMemcpyAsync(d_a[k], A + r0 * n, nr * n * 8, cudaMemcpyHostToDevice, sc));
      CKO(cudaEventRecord(up[k], sc));
      CKO(cudaStreamWaitEvent(sk, up[k], 0));
      for (int q0 = 0; q0 < nq; q0 += twk::kRedQ) {
        twk::rows_apply_multi_kernel<<<(unsigned)((nr + 7) / 8), 256, 0, sk>>>(d_a[k], (long long)n, (int)nr, (int)n, d_X + (size_t)q0 * ldx,
                                                                                ldx, std::min(twk::kRedQ, nq - q0),
                                                                                d_Y + (size_t)q0 * ldy + r0, ldy);
        CKO(cudaGetLastError());
        note_launch();
      }
      CKO(cudaEventRecord(used[k], sk));
    }
    CKO(cudaStreamSynchronize(sk));
    return "";
  };
  std::string err = run();
  for (int i = 0; i < 2; i++) {
    if (d_a[i]) cudaFree(d_a[i]);
    if (up[i]) cudaEventDestroy(up[i]);
    if (used[i]) cudaEventDestroy(used[i]);
  }
  if (sc) cudaStreamDestroy(sc);
  if (sk) cudaStreamDestroy(sk);
  return err;
}

std::string gpu_gram(const double* d_U, long long ldu, int na, const double* d_W, long long ldw, int nb, int n, double* h_G) {
  if (na == 0 || nb == 0) return "";
  DBuf<double> G;
  std::string e;
  if (!(e = G.zeros((size_t)na * nb)).empty()) return e;
  twk::gram_kernel<<<na * nb, 256>>>(d_U, ldu, d_W, ldw, n, nb, G.p);
  CKO(cudaGetLastError());
  note_launch();
  CKO(cudaMemcpy(h_G, G.p, (size_t)na * nb * 8, cudaMemcpyDeviceToHost));
  return "";
}

// tw_reduce_model (thin_wall_solvers.F90:1180-1359), dense-L branch: project the model onto `neigs` basis vectors
// (eig_vec[neigs][nelems], Fortran eig_vec(nelems,neigs)) and write the root-level datasets ThinCurr_reduced reads
// (ThinCurr/_core.py:749-778): ThinCurr_Version, Basis, L, R, Ms, Mc, Msc, Bx/By/Bz, Bx_c/By_c/Bz_c, with the on-disk
// shapes of the reference (Fortran dims reversed).  L V and B V stream the host-resident operators once through the
// device (8 vectors per pass); the k x k / k x n_s Gram products run on the device as well.  The SENSORS/ and COILS/
// metadata groups and the description attributes (:1241-1252,:1265-1285) are not written (root-level writer); nothing
// in ThinCurr_reduced reads them.  B: the reference's dgemm views Bel(:,:,k) -- allocated (nelems,np) -- as an
// (np,nelems) matrix with leading dimension np (:1324-1329), which is the intended product only when np == nelems;
// here Bx(p,q) = sum_e Bel(e,p,1) eig_vec(e,q) is formed, which is what ThinCurr_reduced.reconstruct_Bfield assumes.
std::string reduce_model(Model& m, const Sensors* sens, const std::string& filename, int neigs, const double* eig_vec, bool compute_B) {
  if (neigs <= 0) return "thincurr_reduce_model: no basis vectors";
  if (filename.empty()) return "thincurr_reduce_model: no file name";
  std::string e = need_gpu();
  if (!e.empty()) return e;
  const size_t N = (size_t)m.nelems, K = (size_t)neigs;
  DBuf<double> V, Y;
  std::vector<double> vh(eig_vec, eig_vec + K * N);
  if (!(e = V.up(vh)).empty()) return e;
  if (!(e = Y.zeros(K * std::max(N, (size_t)m.np))).empty()) return e;
  // L: Mat_red(a,b) = V_a . (L V_b), stored [b][a]
  std::vector<double> Lred(K * K), Rred(K * K);
  if (!(e = gpu_host_matrix_multi(m.Lmat.p, N, N, neigs, V.p, (long long)N, Y.p, (long long)N)).empty()) return e;
  if (!(e = gpu_gram(Y.p, (long long)N, neigs, V.p, (long long)N, neigs, (int)N, Lred.data())).empty()) return e;
  // R: sparse apply on the host (O(7 N k)), Gram on the device
  {
    std::vector<double> yr(K * N, 0.0);
    for (size_t q = 0; q < K; q++)
      for (size_t i = 0; i < N; i++) {
        double sacc = 0.0;
        for (int k = m.R_kr[i] - 1; k < m.R_kr[i + 1] - 1; k++) sacc += m.R_val[k] * vh[q * N + (m.R_lc[k] - 1)];
        yr[q * N + i] = sacc;
      }
    CKO(cudaMemcpy(Y.p, yr.data(), K * N * 8, cudaMemcpyHostToDevice));
    if (!(e = gpu_gram(Y.p, (long long)N, neigs, V.p, (long long)N, neigs, (int)N, Rred.data())).empty()) return e;
  }
  std::vector<H5Item> items;
  static const int32_t version = 1;  // tw_idx_ver, thin_wall.F90:160
  items.push_back(H5Item{"ThinCurr_Version", false, {1}, &version});
  items.push_back(H5Item{"Basis", true, {K, N}, eig_vec});
  items.push_back(H5Item{"L", true, {K, K}, Lred.data()});
  items.push_back(H5Item{"R", true, {K, K}, Rred.data()});
  // sensors: Ms(nfloops,neigs) = Ael2sen(nfloops,nelems) V, stored [q][s]
  std::vector<double> Ms, Mc;
  const size_t ns = sens ? sens->floops.size() : 0;
  if (ns > 0) {
    if (!m.Ael2sen.p || (size_t)m.nsensors_built != ns) return "thincurr_reduce_model: sensor mutuals required, but not computed";
    std::vector<double> at(ns * N);
    for (size_t el = 0; el < N; el++)
      for (size_t si = 0; si < ns; si++) at[si * N + el] = m.Ael2sen.p[el * ns + si];
    DBuf<double> A;
    if (!(e = A.up(at)).empty()) return e;
    Ms.resize(K * ns);
    if (!(e = gpu_gram(V.p, (long long)N, neigs, A.p, (long long)N, (int)ns, (int)N, Ms.data())).empty()) return e;
    items.push_back(H5Item{"Ms", true, {K, ns}, Ms.data()});
  }
  if (m.n_icoils > 0) {
    if (!m.Ael2dr.p) return "thincurr_reduce_model: coil mutuals required, but not computed";
    const size_t ni = (size_t)m.n_icoils;
    DBuf<double> A;
    std::vector<double> ah(m.Ael2dr.p, m.Ael2dr.p + ni * N);
    if (!(e = A.up(ah)).empty()) return e;
    Mc.resize(ni * K);
    if (!(e = gpu_gram(A.p, (long long)N, (int)ni, V.p, (long long)N, neigs, (int)N, Mc.data())).empty()) return e;
    items.push_back(H5Item{"Mc", true, {ni, K}, Mc.data()});
    if (ns > 0 && m.Adr2sen.p) items.push_back(H5Item{"Msc", true, {ni, ns}, m.Adr2sen.p});
  }
  std::vector<double> Bred;
  if (compute_B) {
    if (!m.Bel.p && !(e = gpu_bmat(m)).empty()) return e;
    const size_t np = (size_t)m.np;
    Bred.resize(3 * K * np);
    static const char* names[3] = {"Bx", "By", "Bz"};
    static const char* cnames[3] = {"Bx_c", "By_c", "Bz_c"};
    for (int k = 0; k < 3; k++) {
      if (!(e = gpu_host_matrix_multi(m.Bel.p + (size_t)k * np * N, np, N, neigs, V.p, (long long)N, Y.p, (long long)np)).empty()) return e;
      for (size_t q = 0; q < K; q++) CKO(cudaMemcpy(Bred.data() + ((size_t)k * K + q) * np, Y.p + q * np, np * 8, cudaMemcpyDeviceToHost));
      items.push_back(H5Item{names[k], true, {K, np}, Bred.data() + (size_t)k * K * np});
    }
    if (m.n_icoils > 0 && m.Bdr.p)
      for (int k = 0; k < 3; k++)
        items.push_back(H5Item{cnames[k], true, {(uint64_t)m.n_icoils, np}, m.Bdr.p + (size_t)k * m.n_icoils * np});
  }
  return write_h5_file(filename, items);
}

}  // namespace tw
