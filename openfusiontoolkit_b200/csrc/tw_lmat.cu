// tw_lmat.cu -- dense element<->element inductance build on sm_100a (FP64, no tensor cores).
//
// Replaces the O(nc^2) OpenMP loop nest of tw_compute_LmatDirect (src/physics/thin_wall.F90:
// 1008-1126) with an owner-computes tiling.  One persistent CTA per SM pulls output tiles
// (row patch x column patch) from a cost-sorted queue.  For every pair of 64-cell chunks of the
// two patches it
//   A. stages both chunks' SoA geometry records in shared memory (1-D bulk async copies,
//      mbarrier-tracked) and classifies the 4096 cell pairs: the quadrature order of
//      thin_wall.F90:1044-1059 is screened in FP32 on locally shifted coordinates with a rigorous
//      error band and falls back to a bit-exact FP64 evaluation when a threshold is within the band;
//   B. bins the pairs by rule (counting sort in shared memory, bins padded to warp multiples) so
//      that every warp executes ONE rule -- no divergence;
//   C. evaluates T(c1,c2): far pairs (thin_wall.F90:1069-1083) from per-chunk tables of quadrature
//      points held in shared memory ((x,y,z,|x|^2) in a local frame, d^2 = |xi|^2+|xj|^2-2 xi.xj,
//      MUFU.RSQ64H seed + third-order correction = 10 FP64-pipe instructions per 1/r), near pairs
//      (thin_wall.F90:1061-1068) with one half-warp per pair, lanes over the quadrature points of
//      the analytic potential; work is handed out in warp-sized batches from a shared counter;
//   D. contracts T onto the vertex/hole DOFs in two stages (cell x column-DOF partial sums in
//      shared memory, then row-DOF sums) and adds the block into L.
// Every L entry is owned by exactly one CTA and updated by plain read-modify-writes between
// barriers: no atomics on the matrix, deterministic summation.
//
// Role rule (SURVEY hard part 1): for entry (a,b) with a<=b in reference numbering the cell
// carrying `a` is the analytic side of near pairs.  Role-1 values T(c1 analytic) serve entries
// with a<=b, role-2 values T(c2 analytic) entries with a>b; far pairs are role-symmetric.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "quad_tables.h"
#include "tw_device.cuh"
#include "tw_gpu.h"

namespace twk {

using tw::kCH;
using tw::kGeomRows;
#ifndef TW_LMAT_CTAS_PER_SM
#define TW_LMAT_CTAS_PER_SM 1
#endif
constexpr int kCtasPerSM = TW_LMAT_CTAS_PER_SM;
constexpr int NT = 512 / kCtasPerSM;  // threads per CTA
constexpr int NW = NT / 32;
constexpr int CI = kCH / kCtasPerSM;  // row cells per pass: a pass evaluates CI x kCH cell pairs
constexpr int kGeomL = 19;         // geometry rows the L kernel stages (vertices, area, qbasis)
constexpr int TS = kCH + 1;        // row stride of the T tile (bank-conflict-free column access)
constexpr int NCLS = 12;           // far classes 0..6 (iquad 4..10), near classes 7..11 (28,33,46,55,72 points)
constexpr int kListCap = CI * kCH + NCLS * 32;
constexpr int kTabMaxN = kCtasPerSM == 1 ? 25 : 16;  // largest rule served from shared-memory point tables
constexpr int kTabClsMax = kCtasPerSM == 1 ? 6 : 4;   // ... as a far class index
constexpr int kTabPts = kCtasPerSM == 1 ? 34 : 16;    // points the table pool holds (several rules at once)
constexpr int kTabMin = 64;        // fewer pairs of a rule than this: evaluate from the vertices instead
constexpr int UB = 32, US = UB + 1;  // column-DOF block of the contraction and its padded stride

__device__ __constant__ int c_cls_np[NCLS] = {6, 7, 12, 15, 16, 19, 25, 28, 33, 46, 55, 72};
__device__ __forceinline__ int cls_of(int iq) {
  return iq <= 10 ? iq - 4 : (iq == 11 ? 7 : (iq == 12 ? 8 : (iq <= 14 ? 9 : (iq <= 16 ? 10 : 11))));
}

struct LmatArgs {
  // row side / column side patch sets (same pointers for self inductance)
  const tw::ChunkMeta *chunksA, *chunksB;
  const double *geomA, *geomB;
  const int *dminA, *dmaxA, *dminB, *dmaxB;
  const int *chunk_dofA, *chunk_dofB;
  const int *inc_ptrA, *inc_ptrB;
  const uint16_t *incA, *incB;
  const int *patch_chunk_ptrA, *patch_chunk_ptrB;
  const int *dof_origA, *dof_origB;   // internal -> reference DOF id
  const int *row_out;                 // internal row DOF -> output row index or -1
  const tw::Tile* tiles;
  int ntiles;
  int* tile_counter;
  double* out;                        // [rows][ld], column = reference DOF id of the column model
  long long ld;
  double scale;                       // 1/(4 pi)
  int self;                           // 1: self inductance (role rule, mirror), 0: mutual
  int debug_skip;                     // profiling aid: bit0 skip near-field evaluation, bit1 skip far-field evaluation
  unsigned long long* stats;          // [0] far pairs, [1] near T evaluations, [2] 1/r evaluations, [3] phipot evals
};

struct Smem {
  double gI[kGeomL * kCH];
  double gJ[kGeomL * kCH];
  double T[CI * TS];
  union {
    struct {
      double2 tabI[kTabPts * 2 * CI];   // per rule [(p*2+h)*CI + c1]: h=0 (-2x,-2y), h=1 (-2z,|x|^2)
      double2 tabJ[kTabPts * 2 * kCH];  //          [(p*2+h)*kCH + c2]: h=0 (x,y),    h=1 (z,|x|^2)
    } tab;
    double U[3 * CI * US];               // [comp][c1][b] partial sums of the contraction
    struct {
      float vfI[9 * CI], vfJ[9 * kCH];   // vertices in the local frame, FP32 (order screening)
      float flI[CI], flJ[kCH];           // 2 * area
    } scr;
  } w;
  double nI[3 * kCH];   // unit normals (reference formula) of row / column cells
  double nJ[3 * kCH];
  unsigned short list[kListCap];     // pair ids (c1<<6|c2) sorted by class, bins padded with 0xFFFF
  unsigned char iqmap[CI * kCH];     // iquad | need-role-1 << 5 | need-role-2 << 6
  int dminI[kCH], dmaxI[kCH], dminJ[kCH], dmaxJ[kCH];
  int origI[tw::kMaxChunkDof], origJ[tw::kMaxChunkDof];  // reference DOF ids
  int rowI[tw::kMaxChunkDof], rowJ[tw::kMaxChunkDof];    // output rows (or -1)
  int iptrI[tw::kMaxChunkDof + 1], iptrJ[tw::kMaxChunkDof + 1];
  unsigned char hasI[tw::kMaxChunkDof];  // bit h: the DOF has a cell in row pass h of the chunk
  uint16_t incI[tw::kMaxChunkInc], incJ[tw::kMaxChunkInc];
  unsigned long long bar[2];
  int cnt[NCLS], off[NCLS + 1], fill[NCLS];
  int qcls[NCLS + 8], qnb[NCLS + 8], qpt[NCLS + 8];  // work-queue items: class (| 16 = table), batches, table offset
  int gq0[8], gq1[8], gnb[8], ng;                    // table groups: item range and batch count
  int qhead, both_count;
  int tile_id;
};

// ---- far field from the vertices (rules without a table / tiny bins) --------------------------
// T = area_i area_j sum_p sum_q w_p w_q / |x_p(i) - x_q(j)|, same rule on both triangles
// (thin_wall.F90:1069-1083).  j-side points are held in registers in blocks of <= 8; the i-side
// point is recomputed per p.
template <int N>  // @region far_vertex
__device__ __forceinline__ double far_pair(const double* __restrict__ gI, int c1, const double* __restrict__ gJ, int c2,
                                           int iquad) {
  const double* bp = c_qpts + 3 * c_qoff[iquad];
  const double* bw = c_qwts + c_qoff[iquad];
  double Pi[9], Pj[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    Pi[k] = gI[k * kCH + c1];
    Pj[k] = gJ[k * kCH + c2];
  }
  double total = 0.0;
  constexpr int NB = (N + 7) / 8;          // j-side register blocks of <= 8 points
  constexpr int QB = (N + NB - 1) / NB;
#pragma unroll 1
  for (int q0 = 0; q0 < N; q0 += QB) {
    double xj[QB][3], acc[QB];
#pragma unroll
    for (int q = 0; q < QB; q++) {
      const int qq = (q0 + q < N) ? q0 + q : N - 1;  // tail block re-reads the last point (weight masked below)
      double b0 = bp[3 * qq], b1 = bp[3 * qq + 1], b2 = bp[3 * qq + 2];
#pragma unroll
      for (int d = 0; d < 3; d++) xj[q][d] = b0 * Pj[d] + b1 * Pj[3 + d] + b2 * Pj[6 + d];
      acc[q] = 0.0;
    }
#pragma unroll 2
    for (int p = 0; p < N; p++) {
      double a0 = bp[3 * p], a1 = bp[3 * p + 1], a2 = bp[3 * p + 2], wp = bw[p];
      double xi0 = a0 * Pi[0] + a1 * Pi[3] + a2 * Pi[6];
      double xi1 = a0 * Pi[1] + a1 * Pi[4] + a2 * Pi[7];
      double xi2 = a0 * Pi[2] + a1 * Pi[5] + a2 * Pi[8];
#pragma unroll
      for (int q = 0; q < QB; q++) {
        double dx = xi0 - xj[q][0], dy = xi1 - xj[q][1], dz = xi2 - xj[q][2];
        double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
        acc[q] = fma(wp, rsqrt_fast(d2), acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < QB; q++)
      if (q0 + q < N) total = fma(bw[q0 + q], acc[q], total);
  }
  return total * gI[9 * kCH + c1] * gJ[9 * kCH + c2];
}

__device__ __forceinline__ double far_dispatch(const double* gI, int c1, const double* gJ, int c2, int iquad) {  // @region far_dispatch
  switch (iquad) {
    case 4: return far_pair<6>(gI, c1, gJ, c2, iquad);
    case 5: return far_pair<7>(gI, c1, gJ, c2, iquad);
    case 6: return far_pair<12>(gI, c1, gJ, c2, iquad);
    case 7: return far_pair<15>(gI, c1, gJ, c2, iquad);
    case 8: return far_pair<16>(gI, c1, gJ, c2, iquad);
    case 9: return far_pair<19>(gI, c1, gJ, c2, iquad);
    default: return far_pair<25>(gI, c1, gJ, c2, iquad);
  }
}

// ---- far field from the shared-memory point tables ----------------------------------------------
// 10 FP64-pipe instructions per 1/r: 1 add + 3 fma (d^2), 5 (rsqrt correction), 1 fma (weighted sum).
// The QB evaluations of one row point are advanced stage by stage so that QB independent
// dependency chains are in flight (DFMA latency is 8 cycles, the pipe takes one warp every 2).
template <int N, int OFF>  // @region far_tab
__device__ __forceinline__ double far_tab(const double2* __restrict__ tabI, const double2* __restrict__ tabJ, int c1, int c2) {
  // c1: row cell within the pass (stride CI), c2: column cell (stride kCH)
  const double* bw = c_qwts + OFF;  // OFF = TCQ_OFF[iquad]: weights become constant-bank operands
  constexpr int QB = (N == 6) ? 3 : ((N == 15 || N == 25) ? 5 : 4);  // j-side points held in registers
  double total = 0.0;
#pragma unroll 1
  for (int q0 = 0; q0 < N; q0 += QB) {
    double xj[QB], yj[QB], zj[QB], sj[QB], acc[QB];
#pragma unroll
    for (int q = 0; q < QB; q++) {
      const int qq = (q0 + q < N) ? q0 + q : N - 1;
      const double2 u = tabJ[(qq * 2) * kCH + c2], v = tabJ[(qq * 2 + 1) * kCH + c2];
      xj[q] = u.x;
      yj[q] = u.y;
      zj[q] = v.x;
      sj[q] = v.y;
      acc[q] = 0.0;
    }
#pragma unroll(N <= 7 ? N : 2)
    for (int p = 0; p < N; p++) {
      const double2 a = tabI[(p * 2) * CI + c1], b = tabI[(p * 2 + 1) * CI + c1];
      const double wp = bw[p];
      double d2[QB], y0[QB], e[QB], h[QB];
#pragma unroll
      for (int q = 0; q < QB; q++) d2[q] = fma(a.x, xj[q], fma(a.y, yj[q], fma(b.x, zj[q], b.y + sj[q])));
#pragma unroll
      for (int q = 0; q < QB; q++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[q]) : "d"(d2[q]));
#pragma unroll
      for (int q = 0; q < QB; q++) h[q] = d2[q] * y0[q];
#pragma unroll
      for (int q = 0; q < QB; q++) e[q] = fma(-h[q], y0[q], 1.0);
#pragma unroll
      for (int q = 0; q < QB; q++) h[q] = fma(0.375, e[q], 0.5);
#pragma unroll
      for (int q = 0; q < QB; q++) e[q] = e[q] * y0[q];
#pragma unroll
      for (int q = 0; q < QB; q++) y0[q] = fma(e[q], h[q], y0[q]);
#pragma unroll
      for (int q = 0; q < QB; q++) acc[q] = fma(wp, y0[q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < QB; q++)
      if (q0 + q < N) total = fma(bw[q0 + q], acc[q], total);
  }
  return total;
}

__device__ __forceinline__ double far_tab_dispatch(const double2* tabI, const double2* tabJ, int c1, int c2, int cls) {  // @region far_tab_dispatch
  switch (cls) {
    case 0: return far_tab<6, 7>(tabI, tabJ, c1, c2);
    case 1: return far_tab<7, 13>(tabI, tabJ, c1, c2);
    case 2: return far_tab<12, 20>(tabI, tabJ, c1, c2);
    case 3: return far_tab<15, 32>(tabI, tabJ, c1, c2);
    case 4: return far_tab<16, 47>(tabI, tabJ, c1, c2);
    case 5: return far_tab<19, 63>(tabI, tabJ, c1, c2);
    default: return far_tab<25, 82>(tabI, tabJ, c1, c2);
  }
}

// ---- near field ------------------------------------------------------------------------------------
// T = area_q * sum_q w_q phi_{tri A}(x_q(tri Q)) (thin_wall.F90:1061-1068); gA/cA = analytic
// triangle, gQ/cQ = quadrature triangle.  `nl` lanes (16 or 32, aligned group of the warp)
// cooperate on one pair; every lane of the group returns the sum.
__device__ __forceinline__ double near_pair(const double* gA, const double* nA, int cA, const double* gQ, int cQ, int iquad,  // @region near_pair
                                            int gl, int nl, unsigned mask) {
  double PA[9], PQ[9], nh[3];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    PA[k] = gA[k * kCH + cA];
    PQ[k] = gQ[k * kCH + cQ];
  }
  nh[0] = nA[cA];
  nh[1] = nA[kCH + cA];
  nh[2] = nA[2 * kCH + cA];
  const int n = c_qnp[iquad];
  const double* bp = g_qpts + 3 * c_qoff[iquad];  // lane-divergent index -> global copy of the tables
  const double* bw = g_qwts + c_qoff[iquad];
  double s = 0.0;
  for (int q = gl; q < n; q += nl) {
    double b0 = bp[3 * q], b1 = bp[3 * q + 1], b2 = bp[3 * q + 2];
    double x = xquad(b0, b1, b2, PQ[0], PQ[3], PQ[6]);
    double y = xquad(b0, b1, b2, PQ[1], PQ[4], PQ[7]);
    double z = xquad(b0, b1, b2, PQ[2], PQ[5], PQ[8]);
    s += bw[q] * phipot(PA, nh, x, y, z);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(mask, s, o);
  if (nl == 32) s += __shfl_xor_sync(mask, s, 16);
  return s * gQ[9 * kCH + cQ];
}

// ---- order selection: FP32 screen with a rigorous band, exact FP64 fallback -----------------------
// vI/vJ: vertices in a common local frame rounded to FP32 (|v| <= X); delta = bound of the
// coordinate error of a vertex DIFFERENCE (input rounding of both operands, = 2^-23 X * 1.01).
// Returns iquad, or -1 when the decision is not safe in FP32.
__device__ __forceinline__ int iquad_screen(const float* __restrict__ vI, int sI, int c1, const float* __restrict__ vJ, int c2,  // @region iquad_screen
                                            float fl2, float delta) {
  float pi_[9], pj_[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    pi_[k] = vI[k * sI + c1];
    pj_[k] = vJ[k * kCH + c2];
  }
  float d2min = 3.0e38f, d2max = 0.0f;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) {
      float dx = pi_[3 * a] - pj_[3 * b], dy = pi_[3 * a + 1] - pj_[3 * b + 1], dz = pi_[3 * a + 2] - pj_[3 * b + 2];
      float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      d2min = fminf(d2min, d2);
      d2max = fmaxf(d2max, d2);
    }
  // the floor sqrt(2 max(area)) is exact to FP32 rounding; vertex distances carry |err(d)| <= sqrt(3) delta
  const float dmaxv = sqrtf(d2max);
  const bool floor_wins = fl2 >= d2max;
  d2max = fmaxf(d2max, fl2);
  // coincident in FP32 => true dl_min <= sqrt(3) delta; order 18 as soon as that is < 0.3 dl_max
  if (d2min == 0.0f) return (3.0f * delta * delta < 0.09f * d2max) ? 18 : -1;
  // relative error bound of rho^2 = d2min/d2max: 2 err(d)/d per distance + FP32 arithmetic (7 roundings)
  const float e = 1.7320508f * delta;
  float band = 2.0f * e * rsqrtf(d2min) + (floor_wins ? 0.0f : 2.0f * e / dmaxv) + 2.0e-6f;
  band = 1.5f * band + band * band;
  if (!(band < 0.25f)) return -1;
  const float r = d2min / d2max, rlo = r * (1.0f - band), rhi = r * (1.0f + band);
  // candidate from the reference expression in fast FP32, then verified against the exact
  // decision boundaries (iquad >= k <=> rho^2 <= c_thr2f[k-5]) with the band on both sides
  if (rlo > c_thr2f[0]) return 4;
  const float qf = -18.420681f / __logf(1.0f - sqrtf(r));
  int iq = (int)fminf(fmaxf(qf, 4.0f), 18.0f);
  if (iq < 18 && rhi <= c_thr2f[iq - 4]) iq++;           // candidate one too low
  else if (iq > 4 && rlo > c_thr2f[iq - 5]) iq--;         // candidate one too high
  const bool lo_ok = (iq == 4) || (rhi <= c_thr2f[iq - 5]);  // surely iquad >= iq
  const bool hi_ok = (iq == 18) || (rlo > c_thr2f[iq - 4]);  // surely iquad <  iq + 1
  return (lo_ok && hi_ok) ? iq : -1;
}

__device__ __noinline__ int iquad_exact_cells(const double* gI, int c1, const double* gJ, int c2) {  // @region iquad_exact_cells
  double Pi[9], Pj[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    Pi[k] = gI[k * kCH + c1];
    Pj[k] = gJ[k * kCH + c2];
  }
  return iquad_exact(Pi, Pj, 3, 3, fmax(gI[9 * kCH + c1], gJ[9 * kCH + c2]) * 2.0);
}

// (kept for the probes) classification directly in FP64
__device__ __forceinline__ int classify_pair(const double* gI, int c1, const double* gJ, int c2) {  // @region classify_pair
  double Pi[9], Pj[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    Pi[k] = gI[k * kCH + c1];
    Pj[k] = gJ[k * kCH + c2];
  }
  double d2min = 1.e300, d2max = 0.0;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) {
      double dx = Pi[3 * a] - Pj[3 * b], dy = Pi[3 * a + 1] - Pj[3 * b + 1], dz = Pi[3 * a + 2] - Pj[3 * b + 2];
      double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
      d2min = fmin(d2min, d2);
      d2max = fmax(d2max, d2);
    }
  double floor2 = fmax(gI[9 * kCH + c1], gJ[9 * kCH + c2]) * 2.0;
  int iq = iquad_fast(d2min, fmax(d2max, floor2));
  if (iq < 0) iq = iquad_exact(Pi, Pj, 3, 3, floor2);
  return iq;
}

// quadrature-point table of `ncell` cells (first cell c0 of the chunk record g) for one rule in the
// frame centred at (ox,oy,oz); table stride `ts`.  neg=true stores (-2x,-2y),(-2z,|x|^2) (row side),
// else (x,y),(z,|x|^2)
__device__ __forceinline__ void build_table(double2* __restrict__ tab, int ts, const double* __restrict__ g, int c0, int ncell,  // @region build_table
                                            int iquad, int n, double ox, double oy, double oz, bool neg, int tid0, int nthreads) {
  const double* bp = c_qpts + 3 * c_qoff[iquad];
  for (int it = tid0; it < n * ts; it += nthreads) {
    const int p = it / ts, cl = it - p * ts, c = c0 + cl;
    if (cl >= ncell) continue;
    const double b0 = bp[3 * p], b1 = bp[3 * p + 1], b2 = bp[3 * p + 2];
    const double x = (b0 * g[0 * kCH + c] + b1 * g[3 * kCH + c] + b2 * g[6 * kCH + c]) - ox;
    const double y = (b0 * g[1 * kCH + c] + b1 * g[4 * kCH + c] + b2 * g[7 * kCH + c]) - oy;
    const double z = (b0 * g[2 * kCH + c] + b1 * g[5 * kCH + c] + b2 * g[8 * kCH + c]) - oz;
    const double s2 = fma(z, z, fma(y, y, x * x));
    if (neg) {
      tab[(p * 2) * ts + cl] = make_double2(-2.0 * x, -2.0 * y);
      tab[(p * 2 + 1) * ts + cl] = make_double2(-2.0 * z, s2);
    } else {
      tab[(p * 2) * ts + cl] = make_double2(x, y);
      tab[(p * 2 + 1) * ts + cl] = make_double2(z, s2);
    }
  }
}

__device__ __forceinline__ void load_chunk(Smem& S, int side, const LmatArgs& A, int chunk, unsigned long long* bar,  // @region load_chunk
                                           uint32_t& phase) {
  // side 0: row chunk (I), 1: column chunk (J).  Geometry record via one bulk async copy issued
  // by a single thread; the small index lists by all threads; normals computed after arrival.
  const tw::ChunkMeta* cms = side ? A.chunksB : A.chunksA;
  const double* geom = side ? A.geomB : A.geomA;
  double* g = side ? S.gJ : S.gI;
  const tw::ChunkMeta cm = cms[chunk];
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order prior generic accesses before the async write
    mbar_expect_tx(bar, kGeomL * kCH * 8);
    bulk_g2s(g, geom + (size_t)chunk * kGeomRows * kCH, kGeomL * kCH * 8, bar);
  }
  const int* dmn = (side ? A.dminB : A.dminA) + (size_t)chunk * kCH;
  const int* dmx = (side ? A.dmaxB : A.dmaxA) + (size_t)chunk * kCH;
  const int* cdof = (side ? A.chunk_dofB : A.chunk_dofA) + cm.dof_off;
  const int* iptr = (side ? A.inc_ptrB : A.inc_ptrA) + cm.dof_off + chunk;
  const uint16_t* inc = (side ? A.incB : A.incA) + cm.inc_off;
  const int* dorig = side ? A.dof_origB : A.dof_origA;
  int* sdmn = side ? S.dminJ : S.dminI;
  int* sdmx = side ? S.dmaxJ : S.dmaxI;
  int* sorig = side ? S.origJ : S.origI;
  int* srow = side ? S.rowJ : S.rowI;
  int* sptr = side ? S.iptrJ : S.iptrI;
  uint16_t* sinc = side ? S.incJ : S.incI;
  for (int i = threadIdx.x; i < kCH; i += NT) {
    sdmn[i] = dmn[i];
    sdmx[i] = dmx[i];
  }
  for (int i = threadIdx.x; i < cm.ndof; i += NT) {
    const int d = cdof[i];
    sorig[i] = dorig[d];
    // rows of the column side exist only for self inductance (mirror writes)
    srow[i] = (side == 0 || A.self) ? A.row_out[d] : -1;
  }
  for (int i = threadIdx.x; i <= cm.ndof; i += NT) sptr[i] = iptr[i];
  const int ninc = iptr[cm.ndof];
  for (int i = threadIdx.x; i < ninc; i += NT) sinc[i] = inc[i];
  mbar_wait(bar, phase);
  phase ^= 1;
  double* nn = side ? S.nJ : S.nI;
  for (int c = threadIdx.x; c < kCH; c += NT) {
    double P[9], n[3] = {0.0, 0.0, 1.0};
#pragma unroll
    for (int k = 0; k < 9; k++) P[k] = g[k * kCH + c];
    if (c < cm.ncell) tri_normal(P, n);
    nn[c] = n[0];
    nn[kCH + c] = n[1];
    nn[2 * kCH + c] = n[2];
  }
}

// one batch of the work queue: 32 far pairs of one rule (from the vertices) or 2 near pairs.
// List entries are (c1l<<6 | c2) with c1l the row cell within the pass; c1 = cbase + c1l.
__device__ __forceinline__ void run_batch_c0(Smem& S, int cbase, int cls, int first, int lane, bool role2_pass,  // @region run_batch_c0
                                             unsigned long long& st_near, unsigned long long& st_phi) {
  if (cls < 7) {
    const unsigned e = S.list[first + lane];
    if (e != 0xFFFFu) {
      const int c1l = e >> 6, c2 = e & 63;
      S.T[c1l * TS + c2] = far_dispatch(S.gI, cbase + c1l, S.gJ, c2, cls + 4);
    }
  } else {
    const int hw = lane >> 4, hl = lane & 15;
    const unsigned e = S.list[first + hw];
    unsigned m = 0;
    int c1l = 0, c2 = 0, iq = 18;
    if (e != 0xFFFFu) {
      m = S.iqmap[e];
      c1l = e >> 6;
      c2 = e & 63;
      iq = m & 31;
    }
    const bool n1 = m & 32, n2 = m & 64;
    // first pass: role 1 if needed, else role 2; second pass: role 2 of the pairs that need both
    const bool do1 = !role2_pass && n1, do2 = role2_pass ? (n1 && n2) : (!n1 && n2);
    if (do1 || do2) {  // uniform per half-warp; the shuffles name only this half
      const unsigned mask = 0xFFFFu << (16 * hw);
      double v;
      if (do2) v = near_pair(S.gJ, S.nJ, c2, S.gI, cbase + c1l, iq, hl, 16, mask);
      else v = near_pair(S.gI, S.nI, cbase + c1l, S.gJ, c2, iq, hl, 16, mask);
      if (hl == 0) {
        S.T[c1l * TS + c2] = v;
        st_near++;
        st_phi += c_qnp[iq];
      }
    }
  }
}

__global__ void __launch_bounds__(NT, kCtasPerSM) lmat_tile_kernel(const LmatArgs A) {  // @region kernel_head
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phI = 0, phJ = 0;
  unsigned long long st_far = 0, st_near = 0, st_eval = 0, st_phi = 0;
  if (A.stats && tid == 0 && blockIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    A.stats[6] = t0;
  }

  for (;;) {
    if (tid == 0) S.tile_id = atomicAdd(A.tile_counter, 1);
    __syncthreads();
    const int t = S.tile_id;
    if (t >= A.ntiles) break;
    const tw::Tile tile = A.tiles[t];
    const bool diag = tile.flags & 1, mirror = tile.flags & 2;
    const bool want2 = (tile.flags & 4) && A.self;
    const int ci0 = A.patch_chunk_ptrA[tile.pa], ci1 = A.patch_chunk_ptrA[tile.pa + 1];
    const int cj0 = A.patch_chunk_ptrB[tile.pb], cj1 = A.patch_chunk_ptrB[tile.pb + 1];

    for (int ci = ci0; ci < ci1; ci++) {  // @region chunk_loop
      __syncthreads();  // previous contraction finished with the I-side lists
      load_chunk(S, 0, A, ci, &S.bar[0], phI);
      const tw::ChunkMeta cmI = A.chunksA[ci];
      const int ncI = cmI.ncell, ndI = cmI.ndof;
      __syncthreads();
      for (int ia = tid; ia < ndI; ia += NT) {
        unsigned hm = 0;
        for (int i1 = S.iptrI[ia]; i1 < S.iptrI[ia + 1]; i1++) hm |= 1u << ((S.incI[i1] & 63) / CI);
        S.hasI[ia] = (unsigned char)hm;
      }
      for (int cj = cj0; cj < cj1; cj++) {
        __syncthreads();  // previous contraction finished with the J-side lists / T / U
        load_chunk(S, 1, A, cj, &S.bar[1], phJ);
        const tw::ChunkMeta cmJ = A.chunksB[cj];
        const int ncJ = cmJ.ncell, ndJ = cmJ.ndof;
        // local frame: midpoint of the two chunk centres
        const double ox = 0.5 * (cmI.cx + cmJ.cx), oy = 0.5 * (cmI.cy + cmJ.cy), oz = 0.5 * (cmI.cz + cmJ.cz);
        float delta;
        {
          const double hx = 0.5 * (cmI.cx - cmJ.cx), hy = 0.5 * (cmI.cy - cmJ.cy), hz = 0.5 * (cmI.cz - cmJ.cz);
          const double X = sqrt(hx * hx + hy * hy + hz * hz) + fmax(cmI.rad, cmJ.rad);
          delta = (float)(X * 1.21e-7);  // two operands, each rounded to FP32 (2^-24 relative), 1% slack
        }
#pragma unroll 1
        for (int cbase = 0; cbase < ncI; cbase += CI) {  // row cells [cbase, cbase + nI1) of the chunk
          const int nI1 = min(CI, ncI - cbase);
          if (cbase > 0) __syncthreads();  // previous pass finished with T / U / lists
          // ---------------- phase A0: FP32 local-frame vertices, list reset ---------------------------  // @region A0_fp32_stage
          for (int i = tid; i < 9 * kCH; i += NT) {
            const int k = i / kCH, d = k % 3;
            const double o = d == 0 ? ox : (d == 1 ? oy : oz);
            S.w.scr.vfJ[i] = (float)(S.gJ[i] - o);
          }
          for (int i = tid; i < 9 * CI; i += NT) {
            const int k = i / CI, c = i - k * CI, d = k % 3;
            const double o = d == 0 ? ox : (d == 1 ? oy : oz);
            S.w.scr.vfI[i] = (float)(S.gI[k * kCH + cbase + c] - o);
          }
          for (int i = tid; i < kCH; i += NT) S.w.scr.flJ[i] = (float)(2.0 * S.gJ[9 * kCH + i]);
          for (int i = tid; i < CI; i += NT) S.w.scr.flI[i] = (float)(2.0 * S.gI[9 * kCH + cbase + i]);
          for (int i = tid; i < kListCap; i += NT) S.list[i] = 0xFFFFu;
          if (tid < NCLS) {
            S.cnt[tid] = 0;
            S.fill[tid] = 0;
          }
          if (tid == 0) S.both_count = 0;
          __syncthreads();
          // ---------------- phase A: classification ----------------------------------------------------  // @region A_classify
          // warp w handles rows c1l = (w>>1) + (NW/2) m, columns lane + 32 (w&1)
          unsigned mycls = 0;  // 4 bits per iteration: class + 1, 0 = no pair
          {
            const int c2 = lane + 32 * (warp & 1);
#pragma unroll 1
            for (int m = 0; m < CI / (NW / 2); m++) {
              const int c1l = (warp >> 1) + (NW / 2) * m, c1 = cbase + c1l;
              unsigned code = 0;
              int cls = -1;
              if (c1l < nI1 && c2 < ncJ) {
                bool n1, n2 = false;
                if (A.self) {
                  n1 = S.dminI[c1] <= S.dmaxJ[c2];
                  n2 = want2 && !diag && (S.dmaxI[c1] > S.dminJ[c2]);
                } else {
                  n1 = true;
                }
                if (n1 || n2) {
                  int iq = iquad_screen(S.w.scr.vfI, CI, c1l, S.w.scr.vfJ, c2, fmaxf(S.w.scr.flI[c1l], S.w.scr.flJ[c2]), delta);
                  if (iq < 0) iq = iquad_exact_cells(S.gI, c1, S.gJ, c2);
                  code = (unsigned)iq | (n1 ? 32u : 0u) | (n2 ? 64u : 0u);
                  cls = cls_of(iq);
                }
              }
              S.iqmap[c1l * kCH + c2] = (unsigned char)code;  // (T of unused pairs is never read by a used entry)
              const unsigned grp = __match_any_sync(0xffffffffu, cls);
              if (cls >= 0 && lane == __ffs(grp) - 1) atomicAdd(&S.cnt[cls], __popc(grp));
              if (cls >= 7 && (code & 96u) == 96u) atomicAdd(&S.both_count, 1);
              mycls |= (unsigned)(cls + 1) << (4 * m);
            }
          }
          __syncthreads();
          // ---------------- phase B: bin offsets, work queue(s), scatter, point tables -------------------  // @region B_bin
          // Queue items: near classes (largest rules first), far bins too small for a table (evaluated
          // from the vertices), then the table rules by decreasing size.  Table rules are packed into
          // groups whose point tables fit the shared-memory pool together; a group is one barrier
          // interval with one dynamic queue, so a pass normally has a single evaluation phase.
          if (tid == 0) {
            const int order[NCLS] = {11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0};
            int o = 0, nb = 0, nq = 0, ng = 0, pts = 0;
            unsigned long long fp = 0, ev = 0;
            S.gq0[0] = 0;
            for (int k = 0; k < NCLS; k++) {
              const int c = order[k], n = S.cnt[c];
              S.off[c] = o;
              const bool near_c = c >= 7;
              const bool c0 = near_c || c > kTabClsMax || n < kTabMin;  // evaluated analytically / from the vertices
              if (n > 0 && c0) {
                S.qcls[nq] = c;
                S.qnb[nq] = near_c ? (n + 1) / 2 : (n + 31) / 32;
                nb += S.qnb[nq];
                nq++;
              }
              o += near_c ? ((n + 1) & ~1) : ((n + 31) & ~31);
              if (!near_c) {
                fp += n;
                ev += (unsigned long long)n * c_cls_np[c] * c_cls_np[c];
              }
            }
            for (int c = kTabClsMax; c >= 0; c--) {
              const int n = S.cnt[c];
              if (n < kTabMin) continue;
              const int np = c_cls_np[c];
              if (pts + np > kTabPts) {  // close the group
                S.gq1[ng] = nq;
                S.gnb[ng] = nb;
                ng++;
                S.gq0[ng] = nq;
                nb = 0;
                pts = 0;
              }
              S.qcls[nq] = c | 16;  // bit 4: evaluate from the tables
              S.qnb[nq] = (n + 31) / 32;
              S.qpt[nq] = pts;
              nb += S.qnb[nq];
              pts += np;
              nq++;
            }
            S.gq1[ng] = nq;
            S.gnb[ng] = nb;
            S.ng = ng + 1;
            S.qhead = 0;
            st_far += fp;
            st_eval += ev;
          }
          __syncthreads();  // (also: the FP32 scratch is dead, the table region may be written)
          {
#pragma unroll 1
            for (int m = 0; m < CI / (NW / 2); m++) {
              const int cls = (int)((mycls >> (4 * m)) & 15u) - 1;
              const unsigned grp = __match_any_sync(0xffffffffu, cls);
              int base = 0;
              const int leader = __ffs(grp) - 1;
              if (cls >= 0 && lane == leader) base = atomicAdd(&S.fill[cls], __popc(grp));
              base = __shfl_sync(0xffffffffu, base, leader);
              if (cls >= 0) {
                const int c1l = (warp >> 1) + (NW / 2) * m, c2 = lane + 32 * (warp & 1);
                S.list[S.off[cls] + base + __popc(grp & ((1u << lane) - 1u))] = (unsigned short)(c1l * kCH + c2);
              }
            }
          }
          // ---------------- phase C: evaluation, one dynamic queue per table group --------------------------  // @region C_eval
          const int ngroups = S.ng;
#pragma unroll 1
          for (int g = 0; g < ngroups; g++) {
            if (g > 0) {
              __syncthreads();  // previous group is done with the table pool
              if (tid == 0) S.qhead = 0;
            }
            const int q0 = S.gq0[g], q1 = S.gq1[g], total = S.gnb[g];
            for (int k = q0; k < q1; k++) {
              const int qc = S.qcls[k];
              if (!(qc & 16)) continue;
              const int cls = qc & 15, pt = S.qpt[k];
              build_table(S.w.tab.tabI + pt * 2 * CI, CI, S.gI, cbase, nI1, cls + 4, c_cls_np[cls], ox, oy, oz, true, tid, NT);
              build_table(S.w.tab.tabJ + pt * 2 * kCH, kCH, S.gJ, 0, ncJ, cls + 4, c_cls_np[cls], ox, oy, oz, false, tid, NT);
            }
            __syncthreads();
            for (;;) {
              int b = 0;
              if (lane == 0) b = atomicAdd(&S.qhead, 1);
              b = __shfl_sync(0xffffffffu, b, 0);
              if (b >= total) break;
              int k = q0, lb = b;
              while (lb >= S.qnb[k]) {
                lb -= S.qnb[k];
                k++;
              }
              const int qc = S.qcls[k], cls = qc & 15;
              if (A.debug_skip && ((cls >= 7) ? (A.debug_skip & 1) : (A.debug_skip & 2))) continue;
              if (qc & 16) {
                const unsigned e = S.list[S.off[cls] + lb * 32 + lane];
                if (e != 0xFFFFu) {
                  const int c1l = e >> 6, c2 = e & 63, pt = S.qpt[k];
                  S.T[c1l * TS + c2] = far_tab_dispatch(S.w.tab.tabI + pt * 2 * CI, S.w.tab.tabJ + pt * 2 * kCH, c1l, c2, cls) *
                                       S.gI[9 * kCH + cbase + c1l] * S.gJ[9 * kCH + c2];
                }
              } else {
                run_batch_c0(S, cbase, cls, S.off[cls] + lb * (cls >= 7 ? 2 : 32), lane, false, st_near, st_phi);
              }
            }
          }
          // ---------------- phase D: contraction onto DOFs (two passes when both roles are needed) ------  // @region D_contract_ctl
          const bool two_pass = S.both_count > 0;  // written before the phase-A barrier
          for (int pass = 0; pass < (two_pass ? 2 : 1); pass++) {
            __syncthreads();  // T complete (and the table region free for U)
            if (pass == 1) {
              // role-2 values of the near pairs that need both roles
              if (tid == 0) S.qhead = 0;
              __syncthreads();
              int nearb = 0, first_cls_off[5], first_cls_nb[5];
#pragma unroll
              for (int k = 0; k < 5; k++) {
                first_cls_off[k] = S.off[11 - k];
                first_cls_nb[k] = (S.cnt[11 - k] + 1) / 2;
                nearb += first_cls_nb[k];
              }
              for (;;) {
                int b = 0;
                if (lane == 0) b = atomicAdd(&S.qhead, 1);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= nearb) break;
                int k = 0, lb = b;
                while (lb >= first_cls_nb[k]) {
                  lb -= first_cls_nb[k];
                  k++;
                }
                if (A.debug_skip & 1) continue;
                run_batch_c0(S, cbase, 11 - k, first_cls_off[k] + lb * 2, lane, true, st_near, st_phi);
              }
              __syncthreads();
            }
            for (int b0 = 0; b0 < ndJ; b0 += UB) {  // @region D_stage1
              const int nb = min(UB, ndJ - b0);
              if (b0 > 0) __syncthreads();  // stage 2 of the previous block finished with U
              // stage 1: U[c1][b] = sum_{(c2,k2) of b} +-E2[c2][k2] T[c1][c2]; lanes over c1
              for (int it = tid; it < nb * CI; it += NT) {
                const int bl = it / CI, c1l = it - bl * CI;
                double ux = 0.0, uy = 0.0, uz = 0.0;
                const int ib = b0 + bl;
                for (int i2 = S.iptrJ[ib]; i2 < S.iptrJ[ib + 1]; i2++) {
                  const unsigned w2 = S.incJ[i2];
                  const int c2 = w2 & 63, k2 = (w2 >> 6) & 3;
                  double tv = S.T[c1l * TS + c2];
                  if (w2 & 256) tv = -tv;
                  ux = fma(S.gJ[(10 + 3 * k2) * kCH + c2], tv, ux);
                  uy = fma(S.gJ[(11 + 3 * k2) * kCH + c2], tv, uy);
                  uz = fma(S.gJ[(12 + 3 * k2) * kCH + c2], tv, uz);
                }
                S.w.U[(0 * CI + c1l) * US + bl] = ux;
                S.w.U[(1 * CI + c1l) * US + bl] = uy;
                S.w.U[(2 * CI + c1l) * US + bl] = uz;
              }
              __syncthreads();
              // stage 2: L[a][b] += sum_{(c1,k1) of a, c1 in this pass} +-E1[c1][k1] . U[c1][b]; lanes over b.  // @region D_stage2
              // Entries are handled in groups of G: all loads of the old values are issued first, the
              // sums are formed while they are in flight, then the stores (one exposed latency per group).
              constexpr int G = 4;
              for (int it0 = tid; it0 < ndI * UB; it0 += G * NT) {
                double* pa[G];
                double* pb[G];
                double olda[G], oldb[G], accv[G];
#pragma unroll
                for (int g = 0; g < G; g++) {
                  pa[g] = nullptr;
                  pb[g] = nullptr;
                  const int it = it0 + g * NT;
                  if (it >= ndI * UB) continue;
                  const int ia = it >> 5, bl = it & 31;
                  if (bl >= nb) continue;
                  const int ib = b0 + bl;
                  const int oa = S.origI[ia], ob = S.origJ[ib];
                  if (A.self) {
                    const bool role1 = oa <= ob;
                    if (diag && !role1) continue;
                    if (two_pass && role1 != (pass == 0)) continue;
                  }
                  if (!((S.hasI[ia] >> (cbase / CI)) & 1)) continue;  // no cell of this DOF in the pass
                  const int ra = S.rowI[ia];
                  if (ra >= 0) pa[g] = A.out + (long long)ra * A.ld + ob;
                  if (A.self && (mirror || diag) && oa != ob) {
                    const int rb = S.rowJ[ib];
                    if (rb >= 0) pb[g] = A.out + (long long)rb * A.ld + oa;
                  }
                }
#pragma unroll
                for (int g = 0; g < G; g++) {
                  olda[g] = pa[g] ? __ldcg(pa[g]) : 0.0;
                  oldb[g] = pb[g] ? __ldcg(pb[g]) : 0.0;
                }
#pragma unroll
                for (int g = 0; g < G; g++) {
                  accv[g] = 0.0;
                  if (!pa[g] && !pb[g]) continue;
                  const int it = it0 + g * NT;
                  const int ia = it >> 5, bl = it & 31;
                  double acc = 0.0;
                  for (int i1 = S.iptrI[ia]; i1 < S.iptrI[ia + 1]; i1++) {
                    const unsigned w1 = S.incI[i1];
                    const int c1l = (int)(w1 & 63) - cbase, k1 = (w1 >> 6) & 3;
                    if (c1l < 0 || c1l >= CI) continue;
                    const int c1 = w1 & 63;
                    double dsum = S.gI[(10 + 3 * k1) * kCH + c1] * S.w.U[(0 * CI + c1l) * US + bl];
                    dsum = fma(S.gI[(11 + 3 * k1) * kCH + c1], S.w.U[(1 * CI + c1l) * US + bl], dsum);
                    dsum = fma(S.gI[(12 + 3 * k1) * kCH + c1], S.w.U[(2 * CI + c1l) * US + bl], dsum);
                    acc += (w1 & 256) ? -dsum : dsum;
                  }
                  accv[g] = acc * A.scale;
                }
#pragma unroll
                for (int g = 0; g < G; g++) {
                  if (pa[g]) __stcg(pa[g], olda[g] + accv[g]);
                  if (pb[g]) __stcg(pb[g], oldb[g] + accv[g]);
                }
              }
            }
          }
        }
      }
    }
  }
  if (A.stats && tid == 0) {  // load balance: first / last CTA finish time (ns, globaltimer)  // @region tail
    unsigned long long tend;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tend));
    atomicMax(&A.stats[4], tend);
    atomicMin(&A.stats[7], tend);
  }
  if (A.stats) {
    st_near = (unsigned long long)warp_sum((double)st_near);  // counts are < 2^53
    st_phi = (unsigned long long)warp_sum((double)st_phi);
    if (lane == 0) {
      if (st_far) atomicAdd(&A.stats[0], st_far);
      atomicAdd(&A.stats[1], st_near);
      if (st_eval) atomicAdd(&A.stats[2], st_eval);
      atomicAdd(&A.stats[3], st_phi);
    }
  }
}

}  // namespace twk

// =============================================================================================
// host side: device mirrors and launch
// =============================================================================================
namespace tw {

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return std::string(#call) + ": " + cudaGetErrorString(e_);                \
  } while (0)

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(); }

template <class T>
static std::string upload(const std::vector<T>& h, T** d) {
  *d = nullptr;
  size_t n = std::max<size_t>(h.size(), 1);
  CK(cudaMalloc((void**)d, n * sizeof(T)));
  if (!h.empty()) CK(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return "";
}

static double order_threshold(int k) {
  // largest rho with |trunc(ln(1e-8)/ln(1-rho))| >= k, by bisection on the host evaluation of the
  // reference expression (thin_wall.F90:1058)
  auto f = [](double rho) { return std::fabs(std::trunc(std::log(1.0e-8) / std::log(1.0 - rho))); };
  double lo = 0.05, hi = 0.999;  // f(lo) >= 18 >= k, f(hi) < 4
  for (int it = 0; it < 200; it++) {
    double mid = 0.5 * (lo + hi);
    if (mid == lo || mid == hi) break;
    if (f(mid) >= k) lo = mid;
    else hi = mid;
  }
  // walk the last ulps
  while (f(std::nextafter(lo, 1.0)) >= k) lo = std::nextafter(lo, 1.0);
  return lo;
}

std::string gpu_init_constants() {
  static thread_local int done_dev = -1;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (done_dev == dev) return "";
  CK(cudaMemcpyToSymbol(twk::c_qpts, TCQ_PTS, sizeof(TCQ_PTS)));
  CK(cudaMemcpyToSymbol(twk::c_qwts, TCQ_WTS, sizeof(TCQ_WTS)));
  CK(cudaMemcpyToSymbol(twk::c_qnp, TCQ_NP, sizeof(TCQ_NP)));
  CK(cudaMemcpyToSymbol(twk::c_qoff, TCQ_OFF, sizeof(TCQ_OFF)));
  CK(cudaMemcpyToSymbol(twk::g_qpts, TCQ_PTS, sizeof(TCQ_PTS)));
  CK(cudaMemcpyToSymbol(twk::g_qwts, TCQ_WTS, sizeof(TCQ_WTS)));
  double thr[14], thr2[14];
  float thr2f[14];
  for (int k = 5; k <= 18; k++) {
    thr[k - 5] = order_threshold(k);
    thr2[k - 5] = thr[k - 5] * thr[k - 5];
    thr2f[k - 5] = (float)thr2[k - 5];
  }
  CK(cudaMemcpyToSymbol(twk::c_thr2f, thr2f, sizeof(thr2f)));
  CK(cudaMemcpyToSymbol(twk::c_thr, thr, sizeof(thr)));
  CK(cudaMemcpyToSymbol(twk::c_thr2, thr2, sizeof(thr2)));
  CK(cudaFuncSetAttribute(twk::lmat_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(twk::Smem)));
  done_dev = dev;
  return "";
}

std::string DevicePatchSet::upload_from(const PatchSet& ps) {
  release();
  std::string e;
  if (!(e = upload(ps.chunks, &chunks)).empty()) return e;
  if (!(e = upload(ps.geom, &geom)).empty()) return e;
  if (!(e = upload(ps.cell_dmin, &dmin)).empty()) return e;
  if (!(e = upload(ps.cell_dmax, &dmax)).empty()) return e;
  if (!(e = upload(ps.chunk_dof, &chunk_dof)).empty()) return e;
  if (!(e = upload(ps.chunk_inc_ptr, &inc_ptr)).empty()) return e;
  if (!(e = upload(ps.inc, &inc)).empty()) return e;
  if (!(e = upload(ps.patch_chunk_ptr, &patch_chunk_ptr)).empty()) return e;
  if (!(e = upload(ps.dof_orig, &dof_orig)).empty()) return e;
  bytes = ps.chunks.size() * sizeof(ChunkMeta) + ps.geom.size() * 8 + (ps.cell_dmin.size() + ps.cell_dmax.size()) * 4 +
          (ps.chunk_dof.size() + ps.chunk_inc_ptr.size() + ps.patch_chunk_ptr.size() + ps.dof_orig.size()) * 4 + ps.inc.size() * 2;
  return "";
}
void DevicePatchSet::release() {
  cudaFree(chunks);
  cudaFree(geom);
  cudaFree(dmin);
  cudaFree(dmax);
  cudaFree(chunk_dof);
  cudaFree(inc_ptr);
  cudaFree(inc);
  cudaFree(patch_chunk_ptr);
  cudaFree(dof_orig);
  chunks = nullptr;
  geom = nullptr;
  dmin = dmax = chunk_dof = inc_ptr = patch_chunk_ptr = dof_orig = nullptr;
  inc = nullptr;
}

std::string gpu_lmat_tiles(const DevicePatchSet& A, const DevicePatchSet& B, const std::vector<Tile>& tiles,
                           const std::vector<int>& row_out, bool self, double* d_out, long long ld, cudaStream_t stream,
                           unsigned long long* h_stats) {
  std::string e = gpu_init_constants();
  if (!e.empty()) return e;
  if (tiles.empty()) return "";
  Tile* d_tiles = nullptr;
  int* d_row_out = nullptr;
  int* d_counter = nullptr;
  unsigned long long* d_stats = nullptr;
  CK(cudaMalloc((void**)&d_tiles, tiles.size() * sizeof(Tile)));
  CK(cudaMalloc((void**)&d_row_out, std::max<size_t>(row_out.size(), 1) * sizeof(int)));
  CK(cudaMalloc((void**)&d_counter, sizeof(int)));
  CK(cudaMalloc((void**)&d_stats, 8 * sizeof(unsigned long long)));
  CK(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(d_row_out, row_out.data(), row_out.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  CK(cudaMemsetAsync(d_counter, 0, sizeof(int), stream));
  CK(cudaMemsetAsync(d_stats, 0, 8 * sizeof(unsigned long long), stream));
  CK(cudaMemsetAsync(d_stats + 7, 0xff, sizeof(unsigned long long), stream));
  twk::LmatArgs a;
  a.chunksA = A.chunks; a.chunksB = B.chunks;
  a.geomA = A.geom; a.geomB = B.geom;
  a.dminA = A.dmin; a.dmaxA = A.dmax; a.dminB = B.dmin; a.dmaxB = B.dmax;
  a.chunk_dofA = A.chunk_dof; a.chunk_dofB = B.chunk_dof;
  a.inc_ptrA = A.inc_ptr; a.inc_ptrB = B.inc_ptr;
  a.incA = A.inc; a.incB = B.inc;
  a.patch_chunk_ptrA = A.patch_chunk_ptr; a.patch_chunk_ptrB = B.patch_chunk_ptr;
  a.dof_origA = A.dof_orig; a.dof_origB = B.dof_orig;
  a.row_out = d_row_out;
  a.tiles = d_tiles;
  a.ntiles = (int)tiles.size();
  a.tile_counter = d_counter;
  a.out = d_out;
  a.ld = ld;
  a.scale = 1.0 / (4.0 * kPi);
  a.self = self ? 1 : 0;
  a.debug_skip = std::getenv("THINCURR_B200_DEBUG_SKIP") ? std::atoi(std::getenv("THINCURR_B200_DEBUG_SKIP")) : 0;
  a.stats = d_stats;
  int dev = 0, nsm = 148;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  int grid = (int)std::min<size_t>(tiles.size(), (size_t)twk::kCtasPerSM * nsm);
  twk::lmat_tile_kernel<<<grid, twk::NT, sizeof(twk::Smem), stream>>>(a);
  CK(cudaGetLastError());
  note_launch();
  if (h_stats) {
    CK(cudaMemcpyAsync(h_stats, d_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
  }
  // stream-ordered frees keep the call asynchronous
  CK(cudaFreeAsync(d_tiles, stream));
  CK(cudaFreeAsync(d_row_out, stream));
  CK(cudaFreeAsync(d_counter, stream));
  CK(cudaFreeAsync(d_stats, stream));
  return "";
}

}  // namespace tw
