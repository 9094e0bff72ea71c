// tw_lmat.cu -- dense element<->element inductance build on sm_100a (FP64, no tensor cores).
//
// Replaces the O(nc^2) OpenMP loop nest of tw_compute_LmatDirect (src/physics/thin_wall.F90:
// 1008-1126) with an owner-computes tiling.  One persistent CTA per SM pulls output tiles
// (row patch x column patch) from a cost-sorted queue.  One PASS = one pair of 64-cell chunks of the
// two patches (4096 cell pairs), six block barriers:
//   0. staging: a chunk is three bulk async copies (cp.async.bulk ... mbarrier::complete_tx) into a
//      ChunkState slot -- 22 SoA geometry rows, the index record (per-cell min/max DOF, reference DOF ids,
//      CSR incidences) and this launch's output rows; the next column chunk is prefetched during the pass;
//   A. classification: the quadrature order of thin_wall.F90:1044-1059 is screened in FP32 on locally
//      shifted coordinates with a rigorous error band (order 4 from the cells' bounding spheres when they
//      are far enough apart) and falls back to a bit-exact FP64 evaluation when a threshold is within
//      the band;
//   B. binning by rule from per-thread packed histograms + one warp scan (no atomics or votes in the pair
//      loop); lists are row-major, bins padded to warp multiples, every warp executes ONE rule;
//   C. evaluation of T(c1,c2) from a dynamic queue of warp-sized batches: far pairs (thin_wall.F90:1069-1083)
//      from per-chunk tables of quadrature points in shared memory ((x,y,z,|x|^2) in a local frame,
//      d^2 = |xi|^2+|xj|^2-2 xi.xj, MUFU.RSQ64H seed + third-order correction = 10 FP64-pipe instructions
//      per 1/r), near pairs (thin_wall.F90:1061-1068) with one half-warp per pair, lanes over the
//      quadrature points of the analytic potential;
//   D. contraction onto the vertex/hole DOFs: a warp per column DOF forms the pass's 64 x 64 block of
//      contributions in shared memory, then a warp per matrix row adds it into L with lanes along the row.
// Every L entry is owned by exactly one CTA and updated by plain loads and stores between barriers: no
// atomics on the matrix, deterministic summation.  When the rows of both patches of a tile are in the output
// block the transposed entries are left to symmetrize_kernel (thin_wall.F90:1146-1151).
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
constexpr int NT = 512;            // threads per CTA (one persistent CTA per SM; 512 threads x 128 registers)
constexpr int NTC = NT;            // threads that classify pairs (8 per row cell)
constexpr int NW = NT / 32;
constexpr int CI = kCH;            // row cells per pass: a pass evaluates one pair of chunks, CI x kCH cell pairs
constexpr int kGeomL = 22;         // geometry rows the L kernel stages (vertices, area, qbasis, phipot normal)
static_assert(CI == tw::kRowHalf, "row block of the plan records");
constexpr int TS = kCH + 1;        // row stride of the T tile (bank-conflict-free column access)
constexpr int NCLS = 12;           // far classes 0..6 (iquad 4..10), near classes 7..11 (28,33,46,55,72 points)
constexpr int kListCap = CI * kCH + NCLS * 32;
constexpr int kTabMaxN = 25;       // largest rule served from shared-memory point tables
constexpr int kTabClsMax = 6;      // ... as a far class index
constexpr int kTabPts = 34;        // points the table pool holds (several rules at once)
constexpr int kTabMin = 64;        // fewer pairs of a rule than this: evaluate from the vertices instead

__device__ __constant__ int c_cls_np[NCLS] = {6, 7, 12, 15, 16, 19, 25, 28, 33, 46, 55, 72};
__device__ __forceinline__ int cls_of(int iq) {
  return iq <= 10 ? iq - 4 : (iq == 11 ? 7 : (iq == 12 ? 8 : (iq <= 14 ? 9 : (iq <= 16 ? 10 : 11))));
}

struct LmatArgs {
  // row side / column side patch sets (same pointers for self inductance)
  const tw::ChunkMeta *chunksA, *chunksB;
  const double *geomA, *geomB;
  const int *patch_chunk_ptrA, *patch_chunk_ptrB;
  const tw::ChunkAux *auxA, *auxB;    // per-chunk index records
  const int *chunk_row;               // [chunk of A][kMaxChunkDof] output row of each local DOF or -1 (this launch)
  const int *col_map;                 // block builds: reference column DOF id -> output column or -1 (nullptr: identity)
  const tw::Tile* tiles;
  int ntiles;
  int* tile_counter;
  double* out;                        // [rows][ld], column = reference DOF id of the column model
  long long ld;
  double scale;                       // 1/(4 pi)
  int self;                           // 1: self inductance (role rule, mirror), 0: mutual
  int fast_lim;                       // local DOFs per chunk side handled through the shared-memory block (64; 32 or 0 in tests of the direct path)
  int debug_skip;                     // profiling aid: bit0 skip near-field evaluation, bit1 skip far-field evaluation,
                                      // bit2 skip the contraction
  unsigned long long* stats;          // [0] far pairs, [1] near T evaluations, [2] 1/r evaluations, [3] phipot evals
};

// one staged chunk (row or column side): SoA geometry record + index record + output rows, each
// filled by one bulk async copy
struct alignas(16) ChunkState {
  double g[kGeomL * kCH];             // rows 0-8 vertices, 9 area, 10-18 qbasis, 19-21 unit normal (phipot's)
  tw::ChunkAux x;
  int row[tw::kMaxChunkDof];          // output rows (or -1)
  double cx, cy, cz, rad;
  int ncell, ndof;
  unsigned long long bar;             // mbarrier of the bulk copies
};
static_assert(offsetof(ChunkState, x) % 16 == 0 && offsetof(ChunkState, row) % 16 == 0, "bulk copy alignment");

// classification result of one pass (CI x kCH cell pairs): pair lists binned by rule + the work queue
struct PassBuf {
  unsigned short list[kListCap];      // pair ids (c1l<<6|c2) sorted by class, bins padded with 0xFFFF
  unsigned char iqmap[CI * kCH];      // iquad | need-role-1 << 5 | need-role-2 << 6
  int cnt[NCLS], off[NCLS + 1];
  int qcls[NCLS + 8], qnb[NCLS + 8], qpt[NCLS + 8];  // queue items: class (| 16 = table), batches, table offset
  int gq0[8], gq1[8], gnb[8], ng;     // table groups: item range and batch count
  int qhead, both_count;
};

struct Smem {
  ChunkState I;                       // row chunk
  ChunkState J[2];                    // column chunks: the one in use and the next one (prefetched)
  PassBuf pb;
  alignas(16) double T[CI * TS];
  union {                             // phases that never overlap share this region
    struct {
      double2 tabI[kTabPts * 2 * CI];   // per rule [(p*2+h)*CI + c1]: h=0 (-2x,-2y), h=1 (-2z,|x|^2)
      double2 tabJ[kTabPts * 2 * kCH];  //          [(p*2+h)*kCH + c2]: h=0 (x,y),    h=1 (z,|x|^2)
    } tab;
    struct {
      float vfI[9 * CI], vfJ[9 * kCH];  // vertices in the local frame, FP32 (order screening)
      float flI[CI], flJ[kCH];          // 2 * area
      float4 cenI[CI], cenJ[kCH];       // centroid (local frame) and the radius covering the vertices
      char pad0[8192 - (9 * CI + 9 * kCH + CI + kCH) * 4 - (CI + kCH) * 16];
      double E[CI * TS];                // contraction: contribution of the pass to L[row DOF][column DOF] (first 64 x 64)
      char pad1[kTabPts * 2 * CI * 16 - 8192 - CI * TS * 8];  // (the per-warp scratch starts at tabJ)
      double P[NW][3 * CI];             // per warp: products of the contraction
    } w;
  } u;
  int tile_id;
#ifdef TW_LMAT_PROF
  long long prof[16], prof_last[2];
#endif
};
static_assert(sizeof(Smem) <= 232448, "shared memory of one CTA");
template <int N> struct ShowSize;
#ifdef TW_SHOW_SMEM
ShowSize<sizeof(Smem)> show_smem_size;
#endif
#ifdef TW_LMAT_PROF
// section timing of CTA 0 (tuning builds only): who = 0 service thread 0, 1 evaluation thread 0
#define TW_MARK(S, who, t, i)                                   \
  if ((t) == 0 && blockIdx.x == 0) {                            \
    const long long now_ = clock64();                           \
    (S).prof[i] += now_ - (S).prof_last[who];                   \
    (S).prof_last[who] = now_;                                  \
  }
#else
#define TW_MARK(S, who, t, i)
#endif

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

__device__ __noinline__ double far_dispatch(const double* gI, int c1, const double* gJ, int c2, int iquad) {  // @region far_dispatch
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
    double2 an = tabI[c1], bn = tabI[CI + c1];  // row point p+1 is loaded while point p is evaluated
#pragma unroll(N <= 7 ? N : 2)
    for (int p = 0; p < N; p++) {
      const double2 a = an, b = bn;
      if (p + 1 < N) {
        an = tabI[((p + 1) * 2) * CI + c1];
        bn = tabI[((p + 1) * 2 + 1) * CI + c1];
      }
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
__device__ __forceinline__ int iquad_screen(const float (&pi_)[9], const float (&pj_)[9], float fl2, float delta) {  // @region iquad_screen
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

// ---- staging of a chunk (one service thread) ---------------------------------------------------------
// side 0: row chunk, 1: column chunk.  Three bulk async copies (geometry rows, index record, output rows of
// this launch) complete on the slot's mbarrier; nothing is read from global memory by the other threads.
__device__ __forceinline__ void stage_issue(ChunkState& C, int side, const LmatArgs& A, int chunk) {  // @region stage_issue
  const bool rows = side == 0 || A.self;  // rows of the column side exist only for self inductance (mirror writes)
  const uint32_t bytes = kGeomL * kCH * 8 + (uint32_t)sizeof(tw::ChunkAux) + (rows ? (uint32_t)sizeof(C.row) : 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order prior generic accesses before the async writes
  mbar_expect_tx(&C.bar, bytes);
  bulk_g2s(C.g, (side ? A.geomB : A.geomA) + (size_t)chunk * kGeomRows * kCH, kGeomL * kCH * 8, &C.bar);
  bulk_g2s(&C.x, (side ? A.auxB : A.auxA) + chunk, (uint32_t)sizeof(tw::ChunkAux), &C.bar);
  if (rows) bulk_g2s(C.row, A.chunk_row + (size_t)chunk * tw::kMaxChunkDof, (uint32_t)sizeof(C.row), &C.bar);
  const tw::ChunkMeta cm = (side ? A.chunksB : A.chunksA)[chunk];
  C.ncell = cm.ncell;
  C.ndof = cm.ndof;
  C.cx = cm.cx;
  C.cy = cm.cy;
  C.cz = cm.cz;
  C.rad = cm.rad;
}

// ---- prep: classification of the CI x kCH cell pairs of a pass, binned by rule -------------------------
// Stage 0 (before the pass barrier): FP32 local-frame copies of the vertices and the bin counters.
__device__ __forceinline__ void prep_stage(Smem& S, const ChunkState& I, const ChunkState& J, int tid) {  // @region A0_fp32_stage
  PassBuf& pb = S.pb;
  // local frame: midpoint of the two chunk centres
  const double ox = 0.5 * (I.cx + J.cx), oy = 0.5 * (I.cy + J.cy), oz = 0.5 * (I.cz + J.cz);
  for (int i = tid; i < 9 * kCH; i += NT) {
    const int k = i / kCH, dd = k % 3;
    const double o = dd == 0 ? ox : (dd == 1 ? oy : oz);
    S.u.w.vfJ[i] = (float)(J.g[i] - o);
    S.u.w.vfI[i] = (float)(I.g[i] - o);
  }
  for (int i = tid; i < kCH; i += NT) {
    S.u.w.flJ[i] = (float)(2.0 * J.g[9 * kCH + i]);
    S.u.w.flI[i] = (float)(2.0 * I.g[9 * kCH + i]);
  }
  if (tid < 2 * kCH) {  // bounding sphere of every cell (order-4 prefilter of the classification)
    const int c = tid & (kCH - 1);
    const double* g = tid < kCH ? I.g : J.g;
    double P[9];
#pragma unroll
    for (int k = 0; k < 9; k++) P[k] = g[k * kCH + c];
    const double mx = (P[0] + P[3] + P[6]) * (1.0 / 3.0), my = (P[1] + P[4] + P[7]) * (1.0 / 3.0), mz = (P[2] + P[5] + P[8]) * (1.0 / 3.0);
    // radius over the vertices; the dl_max floor sqrt(2 area) <= 1.62 x this radius needs no separate term: the test
    // below implies D - R > 79 R
    double r2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double dx = P[3 * k] - mx, dy = P[3 * k + 1] - my, dz = P[3 * k + 2] - mz;
      r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
    }
    const float4 v = make_float4((float)(mx - ox), (float)(my - oy), (float)(mz - oz), (float)(sqrt(r2) * 1.00001));
    if (tid < kCH) S.u.w.cenI[c] = v;
    else S.u.w.cenJ[c] = v;
  }
  if (tid < NCLS) pb.cnt[tid] = 0;
  if (tid == 0) {
    pb.both_count = 0;
    pb.qhead = 0;
  }
}

// Stage 1 (after the pass barrier).  No atomics or warp votes inside the pair loop: every thread keeps a
// packed histogram of its pairs; one warp scan gives the lane offsets and one shared atomic per (warp, class)
// the warp's share of the bin.  Ends with the bins and the work queue complete (two barriers inside).
__device__ __forceinline__ void prep_classify(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, int flags, int tid,
                                              unsigned long long& st_far, unsigned long long& st_eval) {
  PassBuf& pb = S.pb;
  const int lane = tid & 31, warp = tid >> 5;
  const int ncI = I.ncell, ncJ = J.ncell;
  const bool diag = flags & 1, want2 = (flags & 4) && A.self;
  float delta;
  {
    const double hx = 0.5 * (I.cx - J.cx), hy = 0.5 * (I.cy - J.cy), hz = 0.5 * (I.cz - J.cz);
    const double X = sqrt(hx * hx + hy * hy + hz * hz) + fmax(I.rad, J.rad);
    delta = (float)(X * 1.21e-7);  // two operands, each rounded to FP32 (2^-24 relative), 1% slack
  }
  // the order-4 prefilter pays only where most pairs pass: chunks at least ~25 cell sizes apart (block-uniform)
  bool far_pass;
  {
    const double hx = I.cx - J.cx, hy = I.cy - J.cy, hz = I.cz - J.cz;
    far_pass = sqrt(hx * hx + hy * hy + hz * hz) - I.rad - J.rad > 25.0 * (double)(S.u.w.cenI[0].w + S.u.w.cenJ[0].w);
  }
  // ---------------- classification --------------------------------------------------------------------  // @region A_classify
  // a thread owns one row cell c1 and the kCH/TPR columns c2 = NIT * (t % TPR) + m: the pairs of a thread,
  // and of the TPR threads of a row, are consecutive in a bin, so a batch of 32 pairs of the evaluation
  // reads few distinct row-side table entries (broadcast) and consecutive column-side entries
  constexpr int TPR = NTC / CI, NIT = kCH / TPR;
  static_assert(NIT <= 8 && TPR * CI == NTC && NTC <= NT && NIT * TPR == kCH, "4-bit per-thread class counts");
  unsigned mycls = 0;            // 4 bits per iteration: class + 1, 0 = no pair
  unsigned long long hist = 0;   // 4 bits per class: pairs of this thread
  int nboth = 0;
  const int c1 = tid / TPR, c2b = NIT * (tid % TPR);
  if (tid < NTC && c1 < ncI) {
    float pi_[9];
#pragma unroll
    for (int k = 0; k < 9; k++) pi_[k] = S.u.w.vfI[k * CI + c1];
    const float fli = S.u.w.flI[c1];
    const float4 ci = S.u.w.cenI[c1];
    const int dminI = I.x.dmin[c1], dmaxI = I.x.dmax[c1];
#pragma unroll 2
    for (int m = 0; m < NIT; m++) {
      const int c2 = c2b + m;
      if (c2 < ncJ) {
        bool n1, n2 = false;
        if (A.self) {
          n1 = dminI <= J.x.dmax[c2];
          n2 = want2 && !diag && (dmaxI > J.x.dmin[c2]);
        } else {
          n1 = true;
        }
        if (n1 || n2) {
          float pj_[9];
#pragma unroll
          for (int k = 0; k < 9; k++) pj_[k] = S.u.w.vfJ[k * kCH + c2];
          int iq = -2;
          if (far_pass) {
            // order-4 prefilter.  u = unit vector between the centroids, a_k / b_k = projections of the vertices on u,
            // R = sum of the cells' vertex radii (>= the part of any vertex difference perpendicular to u):
            //   dl_min >= Lb = min b - max a,   dl_max <= Ub = (max b - min a) + R^2 / (2 Lb)
            // (the floor sqrt(2 area) <= 1.62 R is below Ub once D > 10 R), so Lb/Ub > 0.9752 > c_thr[0] => iquad = 4.
            const float4 cj = S.u.w.cenJ[c2];
            float ux = cj.x - ci.x, uy = cj.y - ci.y, uz = cj.z - ci.z;
            const float D2 = fmaf(uz, uz, fmaf(uy, uy, ux * ux)), rr = ci.w + cj.w + 4.0f * delta;
            if (D2 > 100.0f * rr * rr) {
              const float rD = rsqrtf(D2);
              ux *= rD;
              uy *= rD;
              uz *= rD;
              const float a0 = fmaf(pi_[2], uz, fmaf(pi_[1], uy, pi_[0] * ux)), a1 = fmaf(pi_[5], uz, fmaf(pi_[4], uy, pi_[3] * ux)),
                          a2 = fmaf(pi_[8], uz, fmaf(pi_[7], uy, pi_[6] * ux));
              const float b0 = fmaf(pj_[2], uz, fmaf(pj_[1], uy, pj_[0] * ux)), b1 = fmaf(pj_[5], uz, fmaf(pj_[4], uy, pj_[3] * ux)),
                          b2 = fmaf(pj_[8], uz, fmaf(pj_[7], uy, pj_[6] * ux));
              const float lb = fminf(b0, fminf(b1, b2)) - fmaxf(a0, fmaxf(a1, a2)) - 8.0f * delta;
              const float ub = fmaxf(b0, fmaxf(b1, b2)) - fminf(a0, fminf(a1, a2)) + 8.0f * delta;
              if (lb > 0.9752f * fmaf(0.5f * rr, __fdividef(rr, lb), ub)) iq = 4;
            }
          }
          if (iq < 0) {
            iq = iquad_screen(pi_, pj_, fmaxf(fli, S.u.w.flJ[c2]), delta);
            if (iq < 0) iq = iquad_exact_cells(I.g, c1, J.g, c2);
          }
          const int cls = cls_of(iq);
          if (cls >= 7) {  // near pairs carry their order and roles (T of unused pairs is never read by a used entry)
            pb.iqmap[c1 * kCH + c2] = (unsigned char)((unsigned)iq | (n1 ? 32u : 0u) | (n2 ? 64u : 0u));
            nboth += (n1 && n2) ? 1 : 0;
          }
          mycls |= (unsigned)(cls + 1) << (4 * m);
          hist += 1ull << (4 * cls);
        }
      }
    }
  }
  TW_MARK(S, 0, tid, 5)
  // ---------------- bins: warp scan of the packed histograms (10-bit fields, 6 classes per word) -----------  // @region B_bin
  unsigned long long h0 = 0, h1 = 0;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    h0 |= ((hist >> (4 * c)) & 15ull) << (10 * c);
    h1 |= ((hist >> (4 * (c + 6))) & 15ull) << (10 * c);
  }
  unsigned long long s0 = h0, s1 = h1;  // inclusive scan over the lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o);
    if (lane >= o) {
      s0 += t0;
      s1 += t1;
    }
  }
  const unsigned long long w0 = __shfl_sync(0xffffffffu, s0, 31), w1 = __shfl_sync(0xffffffffu, s1, 31);  // warp totals
  s0 -= h0;  // exclusive
  s1 -= h1;
  int wbase = 0;  // lane c < NCLS: offset of this warp's pairs within bin c
  if (lane < NCLS) {
    const int tot = (int)(((lane < 6 ? w0 : w1) >> (10 * (lane % 6))) & 1023ull);
    if (tot) wbase = atomicAdd(&pb.cnt[lane], tot);
  }
  nboth = __reduce_add_sync(0xffffffffu, nboth);
  if (lane == 0 && nboth) atomicAdd(&pb.both_count, nboth);
  __syncthreads();
  TW_MARK(S, 0, tid, 6)
  // bin offsets: near classes (largest rules) first; bins padded to whole batches
  int start = 0;  // lane c < NCLS: first list slot of this warp's pairs of class c
  if (lane < NCLS) {
    int o = 0;
    for (int c = NCLS - 1; c > lane; c--) {
      const int n = pb.cnt[c];
      o += c >= 7 ? ((n + 1) & ~1) : ((n + 31) & ~31);
    }
    start = o + wbase;
    if (warp == 0) {
      const int n = pb.cnt[lane];
      pb.off[lane] = o;
      const int padded = lane >= 7 ? ((n + 1) & ~1) : ((n + 31) & ~31);
      for (int i = n; i < padded; i++) pb.list[o + i] = 0xFFFFu;  // padding of the last batch
    }
  }
  // the work queue, built by the lanes of the last warp (lane c = class c; the grouping of the table rules is computed
  // redundantly by every lane from shuffled counts).  Items: near classes (largest rules first), far bins too small
  // for a table (evaluated from the vertices), then the table rules by decreasing size.  Table rules are packed into
  // groups whose point tables fit the shared-memory pool together; a group is one barrier interval.
  if (warp == NW - 1) {
    const int n = lane < NCLS ? pb.cnt[lane] : 0;
    const bool near_c = lane >= 7;
    const bool is_tab = lane <= kTabClsMax && n >= kTabMin;
    const bool is_c0 = lane < NCLS && n > 0 && !is_tab;
    const int nbat = near_c ? (n + 1) / 2 : (n + 31) / 32;
    const unsigned m_c0 = __ballot_sync(0xffffffffu, is_c0), m_tab = __ballot_sync(0xffffffffu, is_tab);
    const int nq0 = __popc(m_c0);
    if (is_c0) {  // classes in descending order
      const int q = __popc(m_c0 >> (lane + 1));
      pb.qcls[q] = lane;
      pb.qnb[q] = nbat;
    }
    int nb0 = is_c0 ? nbat : 0;  // batches of the vertex / analytic items: all in group 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nb0 += __shfl_xor_sync(0xffffffffu, nb0, o);
    // table rules, descending; every lane replays the packing
    int nq = nq0, ng = 0, pts = 0, nb = nb0;
    int my_q = -1, my_pt = 0;
    int gq0_ = 0;  // first item of the open group
#pragma unroll
    for (int c = kTabClsMax; c >= 0; c--) {
      const int nbc = __shfl_sync(0xffffffffu, nbat, c);
      if (!((m_tab >> c) & 1u)) continue;
      constexpr int npc[7] = {6, 7, 12, 15, 16, 19, 25};
      if (pts + npc[c] > kTabPts) {  // close the group
        if (lane == 0) {
          pb.gq0[ng] = gq0_;
          pb.gq1[ng] = nq;
          pb.gnb[ng] = nb;
        }
        ng++;
        gq0_ = nq;
        nb = 0;
        pts = 0;
      }
      if (lane == c) {
        my_q = nq;
        my_pt = pts;
      }
      nb += nbc;
      pts += npc[c];
      nq++;
    }
    if (lane == 0) {
      pb.gq0[ng] = gq0_;
      pb.gq1[ng] = nq;
      pb.gnb[ng] = nb;
      pb.ng = ng + 1;
    }
    if (my_q >= 0) {
      pb.qcls[my_q] = lane | 16;  // bit 4: evaluate from the tables
      pb.qnb[my_q] = nbat;
      pb.qpt[my_q] = my_pt;
    }
    if (lane < 7) {  // statistics: far pairs and 1/r evaluations of this pass
      constexpr int np2[7] = {36, 49, 144, 225, 256, 361, 625};
      st_far += n;
      st_eval += (unsigned long long)n * np2[lane];
    }
  }
  // ---------------- scatter the pair ids into the bins ------------------------------------------------------  // @region B_scatter
  unsigned long long used = 0;  // 4 bits per class: pairs of this thread already placed
#pragma unroll 1
  for (int m = 0; m < NIT; m++) {
    const int cls = (int)((mycls >> (4 * m)) & 15u) - 1;
    const int st = __shfl_sync(0xffffffffu, start, cls < 0 ? 0 : cls);
    if (cls >= 0) {
      const int excl = (int)(((cls < 6 ? s0 : s1) >> (10 * (cls % 6))) & 1023ull);
      const int mine = (int)((used >> (4 * cls)) & 15ull);
      used += 1ull << (4 * cls);
      pb.list[st + excl + mine] = (unsigned short)(c1 * kCH + c2b + m);
    }
  }
  __syncthreads();  // lists, queue; the FP32 scratch is dead: the table region may be written
}

// ---- evaluation: T(c1,c2) of one pass -----------------------------------------------------------------
// one batch of the work queue: 32 far pairs of one rule (from the vertices) or 2 near pairs.
// List entries are (c1<<6 | c2).
__device__ __forceinline__ void run_batch_c0(const ChunkState& I, const ChunkState& J, const PassBuf& pb, double* __restrict__ T,  // @region run_batch_c0
                                             int cls, int first, int lane, bool role2_pass, unsigned long long& st_near,
                                             unsigned long long& st_phi) {
  if (cls < 7) {
    const unsigned e = pb.list[first + lane];
    if (e != 0xFFFFu) {
      const int c1 = e >> 6, c2 = e & 63;
      T[c1 * TS + c2] = far_dispatch(I.g, c1, J.g, c2, cls + 4);
    }
  } else {
    const int hw = lane >> 4, hl = lane & 15;
    const unsigned e = pb.list[first + hw];
    unsigned m = 0;
    int c1 = 0, c2 = 0, iq = 18;
    if (e != 0xFFFFu) {
      m = pb.iqmap[e];
      c1 = e >> 6;
      c2 = e & 63;
      iq = m & 31;
    }
    const bool n1 = m & 32, n2 = m & 64;
    // first pass: role 1 if needed, else role 2; second pass: role 2 of the pairs that need both
    const bool do1 = !role2_pass && n1, do2 = role2_pass ? (n1 && n2) : (!n1 && n2);
    if (do1 || do2) {  // uniform per half-warp; the shuffles name only this half
      const unsigned mask = 0xFFFFu << (16 * hw);
      double v;
      if (do2) v = near_pair(J.g, J.g + 19 * kCH, c2, I.g, c1, iq, hl, 16, mask);
      else v = near_pair(I.g, I.g + 19 * kCH, c1, J.g, c2, iq, hl, 16, mask);
      if (hl == 0) {
        T[c1 * TS + c2] = v;
        st_near++;
        st_phi += c_qnp[iq];
      }
    }
  }
}

__device__ __forceinline__ void eval_pass(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, int tid,
                                          unsigned long long& st_near, unsigned long long& st_phi) {
  PassBuf& pb = S.pb;
  double* __restrict__ T = S.T;
  const int lane = tid & 31;
  const int ncI = I.ncell, ncJ = J.ncell;
  const double ox = 0.5 * (I.cx + J.cx), oy = 0.5 * (I.cy + J.cy), oz = 0.5 * (I.cz + J.cz);
  const int ngroups = pb.ng;
#pragma unroll 1
  for (int g = 0; g < ngroups; g++) {  // @region C_eval
    if (g > 0) {
      __syncthreads();  // previous group is done with the table pool and the queue head
      if (tid == 0) pb.qhead = 0;
    }
    const int q0 = pb.gq0[g], q1 = pb.gq1[g], total = pb.gnb[g];
    bool any_tab = false;
    for (int k = q0; k < q1; k++) {
      const int qc = pb.qcls[k];
      if (!(qc & 16)) continue;
      any_tab = true;
      const int cls = qc & 15, pt = pb.qpt[k];
      build_table(S.u.tab.tabI + pt * 2 * CI, CI, I.g, 0, ncI, cls + 4, c_cls_np[cls], ox, oy, oz, true, tid, NT);
      build_table(S.u.tab.tabJ + pt * 2 * kCH, kCH, J.g, 0, ncJ, cls + 4, c_cls_np[cls], ox, oy, oz, false, tid, NT);
    }
    if (any_tab || g > 0) __syncthreads();
    // dynamic queue; the next batch index is fetched while the current batch is evaluated
    int bnext = 0;
    if (lane == 0) bnext = atomicAdd(&pb.qhead, 1);
    for (;;) {
      const int b = __shfl_sync(0xffffffffu, bnext, 0);
      if (b >= total) break;
      if (lane == 0) bnext = atomicAdd(&pb.qhead, 1);
      int k = q0, lb = b;
      while (lb >= pb.qnb[k]) {
        lb -= pb.qnb[k];
        k++;
      }
      const int qc = pb.qcls[k], cls = qc & 15;
      if (A.debug_skip && ((cls >= 7) ? (A.debug_skip & 1) : (A.debug_skip & 2))) continue;
      if (qc & 16) {
        const unsigned e = pb.list[pb.off[cls] + lb * 32 + lane];
        if (e != 0xFFFFu) {
          const int c1 = e >> 6, c2 = e & 63, pt = pb.qpt[k];
          T[c1 * TS + c2] = far_tab_dispatch(S.u.tab.tabI + pt * 2 * CI, S.u.tab.tabJ + pt * 2 * kCH, c1, c2, cls) * I.g[9 * kCH + c1] *
                            J.g[9 * kCH + c2];
        }
      } else {
        run_batch_c0(I, J, pb, T, cls, pb.off[cls] + lb * (cls >= 7 ? 2 : 32), lane, false, st_near, st_phi);
      }
    }
  }
}

// second role: re-evaluate the near pairs that need both roles (all other T values stay)
__device__ __forceinline__ void eval_role2(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, int tid,  // @region C_role2
                                           unsigned long long& st_near, unsigned long long& st_phi) {
  PassBuf& pb = S.pb;
  const int lane = tid & 31;
  int nearb = 0, first_cls_off[5], first_cls_nb[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    first_cls_off[k] = pb.off[11 - k];
    first_cls_nb[k] = (pb.cnt[11 - k] + 1) / 2;
    nearb += first_cls_nb[k];
  }
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(&pb.qhead, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= nearb) break;
    int k = 0, lb = b;
    while (lb >= first_cls_nb[k]) {
      lb -= first_cls_nb[k];
      k++;
    }
    if (A.debug_skip & 1) continue;
    run_batch_c0(I, J, pb, S.T, 11 - k, first_cls_off[k] + lb * 2, lane, true, st_near, st_phi);
  }
}

// ---- drain: contraction of T onto the DOFs, added into L ------------------------------------------------
// A warp takes column DOFs b = warp, warp+16, ... in blocks of DB.  For a block the old values of all its
// entries (row DOFs a = lane, lane+32; entry (a,b) and its mirror) are loaded first, so one memory latency is
// paid per block.  Then per b -- stage 1, lanes = row cells c1 (two per lane):
// u(c1) = sum_{(c2,k2) of b} +-q2[c2][k2] T[c1][c2]; the three products q1[c1][k].u(c1) (cell c1's contribution
// to its vertices) go to a per-warp scratch; stage 2, lanes = row DOFs a: sum of the scratch entries of a's
// cells, added to the old value and stored.  Every entry is owned by this CTA: plain loads and stores, no
// atomics, no block-wide barriers.
// role_sel 0: all entries; 1: only entries with a <= b (a second-role pass follows); 2: only a > b
struct DrainSel {
  const LmatArgs* A;
  bool diag, mirror;  // mirror: write the transposed entry too
  int role_sel;
  // output addresses of entry (ia, ib) and of its mirror; false if the entry is not written in this pass
  __device__ __forceinline__ bool addr(const ChunkState& I, const ChunkState& J, int ia, int ib, double*& pa, double*& pm) const {
    pa = nullptr;
    pm = nullptr;
    const int oa = I.x.orig[ia], ob = J.x.orig[ib];
    if (A->self) {
      const bool role1 = oa <= ob;
      if ((diag && !role1) || (role_sel == 1 && !role1) || (role_sel == 2 && role1)) return false;
    }
    const int ra = I.row[ia];
    const int oc = A->col_map ? A->col_map[ob] : ob;
    if (ra >= 0 && oc >= 0) pa = A->out + (long long)ra * A->ld + oc;
    if (A->self && mirror && oa != ob) {
      const int rb = J.row[ib];
      if (rb >= 0) pm = A->out + (long long)rb * A->ld + oa;
    }
    return pa || pm;
  }
};

__device__ __forceinline__ void drain_pass(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, int flags, int role_sel,  // @region D_drain
                                           int tid) {
  const double* __restrict__ T = S.T;
  DrainSel sel;
  sel.A = &A;
  sel.diag = flags & 1;
  sel.mirror = (flags & 3) && !(flags & 8);
  sel.role_sel = role_sel;
  const int ndI = I.ndof, ndJ = J.ndof;
  const int lane = tid & 31, warp = tid >> 5;
  double* __restrict__ P = S.u.w.P[warp];  // [3][CI] products of this warp
  double* __restrict__ E = S.u.w.E;        // [row DOF < 64][column DOF < 64] contribution of this pass, stride TS
  const int lim = A.fast_lim;              // 64 (tests of the direct-write path: 32 or 0)
  // ---- D1: contributions of the pass.  This lane's row DOFs a = lane, lane+32 with their incidence lists (scratch
  // index k*64+cell = low 8 bits of the incidence code, bit 8 = negative) packed in registers (longer lists: slow path)
  constexpr int MI = 8;
  int ninc[2];
  unsigned codes[2][MI / 2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int ia = lane + 32 * r;
    ninc[r] = 0;
#pragma unroll
    for (int i = 0; i < MI / 2; i++) codes[r][i] = 0;
    if (ia < ndI) {
      const int i0 = I.x.iptr[ia];
      ninc[r] = I.x.iptr[ia + 1] - i0;
#pragma unroll
      for (int i = 0; i < MI; i++)
        if (i < ninc[r]) codes[r][i >> 1] |= (unsigned)I.x.inc[i0 + i] << (16 * (i & 1));
    }
  }
#pragma unroll 1
  for (int ib = warp; ib < ndJ; ib += NW) {
    // stage 1
    double ux0 = 0.0, uy0 = 0.0, uz0 = 0.0, ux1 = 0.0, uy1 = 0.0, uz1 = 0.0;
    const int i2e = J.x.iptr[ib + 1];
#pragma unroll 4
    for (int i2 = J.x.iptr[ib]; i2 < i2e; i2++) {
      const unsigned w2 = J.x.inc[i2];
      const int c2 = w2 & 63, k2 = (w2 >> 6) & 3;
      double t0 = T[lane * TS + c2], t1 = T[(lane + 32) * TS + c2];
      if (w2 & 256) {
        t0 = -t0;
        t1 = -t1;
      }
      const double qx = J.g[(10 + 3 * k2) * kCH + c2], qy = J.g[(11 + 3 * k2) * kCH + c2], qz = J.g[(12 + 3 * k2) * kCH + c2];
      ux0 = fma(qx, t0, ux0);
      uy0 = fma(qy, t0, uy0);
      uz0 = fma(qz, t0, uz0);
      ux1 = fma(qx, t1, ux1);
      uy1 = fma(qy, t1, uy1);
      uz1 = fma(qz, t1, uz1);
    }
    __syncwarp();  // stage 2 of the previous column DOF is done with the scratch
#pragma unroll
    for (int k = 0; k < 3; k++) {
      P[k * CI + lane] = fma(I.g[(12 + 3 * k) * kCH + lane], uz0, fma(I.g[(11 + 3 * k) * kCH + lane], uy0, I.g[(10 + 3 * k) * kCH + lane] * ux0));
      P[k * CI + lane + 32] =
          fma(I.g[(12 + 3 * k) * kCH + lane + 32], uz1, fma(I.g[(11 + 3 * k) * kCH + lane + 32], uy1, I.g[(10 + 3 * k) * kCH + lane + 32] * ux1));
    }
    __syncwarp();
    // stage 2
    const bool fast_b = ib < lim;
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int ia = lane + 32 * r;
      if (ia >= ndI || 32 * r >= lim) continue;
      double acc = 0.0;
      if (ninc[r] <= MI) {
#pragma unroll
        for (int i = 0; i < MI; i++) {
          const unsigned code = (codes[r][i >> 1] >> (16 * (i & 1))) & 0x1FFu;
          if (i < ninc[r]) {
            const double v = P[code & 255u];
            acc += (code & 256u) ? -v : v;
          }
        }
      } else {
        for (int i1 = I.x.iptr[ia]; i1 < I.x.iptr[ia + 1]; i1++) {
          const unsigned w1 = I.x.inc[i1];
          const double v = P[w1 & 255u];
          acc += (w1 & 256u) ? -v : v;
        }
      }
      if (fast_b) {
        E[ia * TS + ib] = acc;
      } else {  // more than 64 column DOFs in the chunk (rare): write directly
        double *pa, *pm;
        if (sel.addr(I, J, ia, ib, pa, pm)) {
          if (pa) __stcg(pa, fma(acc, A.scale, __ldcg(pa)));  // (every path: one fused scale-and-add)
          if (pm) __stcg(pm, fma(acc, A.scale, __ldcg(pm)));
        }
      }
    }
    // more than 64 row DOFs in the chunk (rare): write directly
#pragma unroll 1
    for (int ia = lane + lim; ia < ndI; ia += 32) {
      double *pa, *pm;
      if (!sel.addr(I, J, ia, ib, pa, pm)) continue;
      double acc = 0.0;
      for (int i1 = I.x.iptr[ia]; i1 < I.x.iptr[ia + 1]; i1++) {
        const unsigned w1 = I.x.inc[i1];
        const double v = P[w1 & 255u];
        acc += (w1 & 256u) ? -v : v;
      }
      if (pa) __stcg(pa, fma(acc, A.scale, __ldcg(pa)));  // (every path: one fused scale-and-add)
      if (pm) __stcg(pm, fma(acc, A.scale, __ldcg(pm)));
    }
  }
  __syncthreads();  // the block E is complete
  TW_MARK(S, 0, tid, 7)
  // ---- D2: add the block into L.  A warp takes rows a = warp, warp+16, ...; lanes run over the column DOFs, whose
  // reference ids ascend within a chunk: neighbouring lanes touch neighbouring addresses of one matrix row.  The old
  // values of all the warp's entries are loaded first (one memory latency), then added and stored.  Every entry is
  // owned by this CTA: plain loads and stores, no atomics.
  const int nbJ = min(ndJ, lim), naI = min(ndI, lim);
  {
    int ob[2], oc[2];  // reference id (role rule) and output column of this lane's column DOFs
    ob[0] = lane < nbJ ? J.x.orig[lane] : 0;
    ob[1] = lane + 32 < nbJ ? J.x.orig[lane + 32] : 0;
    oc[0] = A.col_map ? (lane < nbJ ? A.col_map[ob[0]] : -1) : ob[0];
    oc[1] = A.col_map ? (lane + 32 < nbJ ? A.col_map[ob[1]] : -1) : ob[1];
    constexpr int RW = kCH / NW;  // rows per warp
    double old[RW][2];
    unsigned use = 0;
#pragma unroll
    for (int q = 0; q < RW; q++) {
      const int ia = warp + NW * q;
      double* rowp = nullptr;
      int oa = 0;
      if (ia < naI) {
        const int ra = I.row[ia];
        oa = I.x.orig[ia];
        if (ra >= 0) rowp = A.out + (long long)ra * A.ld;
      }
#pragma unroll
      for (int hb = 0; hb < 2; hb++) {
        old[q][hb] = 0.0;
        bool ok = rowp != nullptr && lane + 32 * hb < nbJ && oc[hb] >= 0;
        if (ok && A.self) {
          const bool role1 = oa <= ob[hb];
          if ((sel.diag && !role1) || (role_sel == 1 && !role1) || (role_sel == 2 && role1)) ok = false;
        }
        if (ok) {
          use |= 1u << (2 * q + hb);
          old[q][hb] = __ldcg(rowp + oc[hb]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < RW; q++) {
      const int ia = warp + NW * q;
      if (ia >= naI) continue;
      const int ra = I.row[ia];
      double* rowp = A.out + (long long)(ra >= 0 ? ra : 0) * A.ld;
#pragma unroll
      for (int hb = 0; hb < 2; hb++)
        if ((use >> (2 * q + hb)) & 1u) __stcg(rowp + oc[hb], fma(E[ia * TS + lane + 32 * hb], A.scale, old[q][hb]));
    }
  }
  if (sel.mirror) {  // transposed entries (rows of the column patch are in the output block, rows of this patch are not)
    int oa[2];
    oa[0] = lane < naI ? I.x.orig[lane] : 0;
    oa[1] = lane + 32 < naI ? I.x.orig[lane + 32] : 0;
#pragma unroll 1
    for (int ib = warp; ib < nbJ; ib += NW) {
      const int rb = J.row[ib];
      if (rb < 0) continue;
      const int ob = J.x.orig[ib];
      double* rowp = A.out + (long long)rb * A.ld;
#pragma unroll
      for (int ha = 0; ha < 2; ha++) {
        const int ia = lane + 32 * ha;
        if (ia >= naI || oa[ha] == ob) continue;
        if (A.self) {
          const bool role1 = oa[ha] <= ob;
          if ((sel.diag && !role1) || (role_sel == 1 && !role1) || (role_sel == 2 && role1)) continue;
        }
        __stcg(rowp + oa[ha], fma(E[ia * TS + ib], A.scale, __ldcg(rowp + oa[ha])));
      }
    }
  }
  TW_MARK(S, 0, tid, 8)
}

// ---- the kernel -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) lmat_tile_kernel(const LmatArgs A) {  // @region kernel_head
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&S.I.bar, 1);
    mbar_init(&S.J[0].bar, 1);
    mbar_init(&S.J[1].bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#ifdef TW_LMAT_PROF
  if (tid < 16) S.prof[tid] = 0;
  if (tid < 2) S.prof_last[tid] = clock64();
#endif
  __syncthreads();
  uint32_t phI = 0, phJ = 0;  // mbarrier phases: row slot, bit s = column slot s
  unsigned long long st_far = 0, st_near = 0, st_eval = 0, st_phi = 0;
  if (A.stats && tid == 0 && blockIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    A.stats[6] = t0;
  }

  for (;;) {
    __syncthreads();  // the previous tile is drained (chunk slots, T and the scratch are free)
    if (tid == 0) S.tile_id = atomicAdd(A.tile_counter, 1);
    __syncthreads();
    const int t = S.tile_id;
    if (t >= A.ntiles) break;
    const tw::Tile tile = A.tiles[t];
    const int flags = tile.flags;
    const int ci0 = A.patch_chunk_ptrA[tile.pa], ci1 = A.patch_chunk_ptrA[tile.pa + 1];
    const int cj0 = A.patch_chunk_ptrB[tile.pb], cj1 = A.patch_chunk_ptrB[tile.pb + 1];
    if (ci0 >= ci1 || cj0 >= cj1) continue;
    int js = 0;  // column slot of the current pass
    if (tid == 0) {
      stage_issue(S.I, 0, A, ci0);
      stage_issue(S.J[0], 1, A, cj0);
    }
    __syncthreads();  // chunk headers written by the staging thread
    mbar_wait(&S.I.bar, phI);
    phI ^= 1;
    TW_MARK(S, 0, tid, 0)
    for (int ci = ci0; ci < ci1; ci++) {  // @region chunk_loop
      for (int cj = cj0; cj < cj1; cj++) {
        const ChunkState& I = S.I;
        const ChunkState& J = S.J[js];
        mbar_wait(&S.J[js].bar, (phJ >> js) & 1u);
        phJ ^= 1u << js;
        prep_stage(S, I, J, tid);
        __syncthreads();  // pass barrier: FP32 copies ready; everybody is done with the previous pass
        TW_MARK(S, 0, tid, 1)
        // prefetch the next column chunk (its slot was used by the previous pass)
        const bool lastj = cj + 1 == cj1, lasti = ci + 1 == ci1;
        if (tid == 0 && !(lastj && lasti)) stage_issue(S.J[js ^ 1], 1, A, lastj ? cj0 : cj + 1);
        prep_classify(S, A, I, J, flags, tid, st_far, st_eval);
        TW_MARK(S, 0, tid, 2)
        const bool two_pass = S.pb.both_count > 0;
        eval_pass(S, A, I, J, tid, st_near, st_phi);
        __syncthreads();  // T complete; the table region is free for the contraction scratch
        TW_MARK(S, 0, tid, 3)
        if (!(A.debug_skip & 4)) drain_pass(S, A, I, J, flags, two_pass ? 1 : 0, tid);
        if (two_pass) {
          if (tid == 0) S.pb.qhead = 0;
          __syncthreads();  // first contraction done with T
          eval_role2(S, A, I, J, tid, st_near, st_phi);
          __syncthreads();
          if (!(A.debug_skip & 4)) drain_pass(S, A, I, J, flags, 2, tid);
        }
        TW_MARK(S, 0, tid, 4)
        js ^= 1;
      }
      if (ci + 1 < ci1) {
        __syncthreads();  // everybody is done with the row chunk
        if (tid == 0) stage_issue(S.I, 0, A, ci + 1);
        __syncthreads();
        mbar_wait(&S.I.bar, phI);
        phI ^= 1;
      }
    }
  }
#ifdef TW_LMAT_PROF
  __syncthreads();
  if (A.stats && blockIdx.x == 0 && tid < 16) A.stats[8 + tid] = (unsigned long long)S.prof[tid];
#endif
  if (A.stats && tid == 0) {  // load balance: first / last CTA finish time (ns, globaltimer)  // @region tail
    unsigned long long tend;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tend));
    atomicMax(&A.stats[4], tend);
    atomicMin(&A.stats[7], tend);
  }
  if (A.stats) {
    st_near = (unsigned long long)warp_sum((double)st_near);  // counts are < 2^53
    st_phi = (unsigned long long)warp_sum((double)st_phi);
    st_far = (unsigned long long)warp_sum((double)st_far);
    st_eval = (unsigned long long)warp_sum((double)st_eval);
    if (lane == 0) {
      if (st_far) atomicAdd(&A.stats[0], st_far);
      if (st_near) atomicAdd(&A.stats[1], st_near);
      if (st_eval) atomicAdd(&A.stats[2], st_eval);
      if (st_phi) atomicAdd(&A.stats[3], st_phi);
    }
  }
}

// Symmetrisation of the owned diagonal block: for internal DOFs i, j in [i0, i1) the tile kernel computed entry
// [i][j] iff patch(i) < patch(j), or the patches are equal and orig(i) <= orig(j); the other one is its copy
// (thin_wall.F90:1146-1151).  32 x 32 blocks of internal DOFs through shared memory: internal order is spatial, so
// the reference ids of a block are a few short runs and both the reads (rows j, columns orig(i)) and the writes
// (rows i, columns orig(j)) touch a few sectors per row instead of one per element.
__global__ void symmetrize_kernel(int i0, int i1, const int* __restrict__ dof_orig, const int* __restrict__ dof_patch,
                                  double* __restrict__ out, long long ld) {
  __shared__ double tile[32][33];
  __shared__ int oi_s[32], pi_s[32], oj_s[32], pj_s[32];
  const int ib = i0 + 32 * blockIdx.x, jb = i0 + 32 * blockIdx.y;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (ty == 0) {
    const int i = ib + tx;
    oi_s[tx] = i < i1 ? dof_orig[i] : -1;
    pi_s[tx] = i < i1 ? dof_patch[i] : -1;
  } else if (ty == 1) {
    const int j = jb + tx;
    oj_s[tx] = j < i1 ? dof_orig[j] : -1;
    pj_s[tx] = j < i1 ? dof_patch[j] : -1;
  }
  __syncthreads();
  // does any entry of this block pair need a copy?  (block-uniform early exit: patches are contiguous DOF ranges)
  if (pj_s[0] > pi_s[min(31, i1 - 1 - ib)]) return;
  for (int jj = ty; jj < 32; jj += 8) {  // read [j][orig(i)], lanes over i
    const int j = jb + jj;
    if (j < i1 && oi_s[tx] >= 0) tile[jj][tx] = __ldcg(out + (long long)(j - i0) * ld + oi_s[tx]);
  }
  __syncthreads();
  for (int ii = ty; ii < 32; ii += 8) {  // write [i][orig(j)], lanes over j
    const int i = ib + ii;
    if (i >= i1 || oj_s[tx] < 0) continue;
    const int pi = pi_s[ii], oi = oi_s[ii], pj = pj_s[tx], oj = oj_s[tx];
    if (pj < pi || (pj == pi && oj < oi)) out[(long long)(i - i0) * ld + oj] = tile[tx][ii];
  }
}

// Transposed copy between the row blocks of two shards (possibly on two devices with peer access): for internal DOFs
// i in [i0,i1) (rows of dst, row index i-i0) and j in [j0,j1) (rows of src, row index j-j0, an earlier shard that
// computed every [j][orig(i)]): dst[i][orig(j)] = src[j][orig(i)]  (thin_wall.F90:1146-1151 across shards).
__global__ void symmetrize_cross_kernel(int i0, int i1, int j0, int j1, const int* __restrict__ dof_orig, double* __restrict__ dst,
                                        const double* __restrict__ src, long long ld) {
  __shared__ double tile[32][33];
  __shared__ int oi_s[32], oj_s[32];
  const int ib = i0 + 32 * blockIdx.x, jb = j0 + 32 * blockIdx.y;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (ty == 0) oi_s[tx] = ib + tx < i1 ? dof_orig[ib + tx] : -1;
  else if (ty == 1) oj_s[tx] = jb + tx < j1 ? dof_orig[jb + tx] : -1;
  __syncthreads();
  for (int jj = ty; jj < 32; jj += 8) {  // read src[j][orig(i)], lanes over i
    const int j = jb + jj;
    if (j < j1 && oi_s[tx] >= 0) tile[jj][tx] = __ldcg(src + (long long)(j - j0) * ld + oi_s[tx]);
  }
  __syncthreads();
  for (int ii = ty; ii < 32; ii += 8) {  // write dst[i][orig(j)], lanes over j
    const int i = ib + ii;
    if (i < i1 && oj_s[tx] >= 0) dst[(long long)(i - i0) * ld + oj_s[tx]] = tile[tx][ii];
  }
}

// output row of every local DOF of every chunk for this launch (row_out: internal DOF -> row or -1)
__global__ void chunk_rows_kernel(int nchunk, const tw::ChunkMeta* __restrict__ chunks, const int* __restrict__ chunk_dof,
                                  const int* __restrict__ row_out, int* __restrict__ chunk_row) {
  const int ch = blockIdx.x, i = threadIdx.x;
  if (ch >= nchunk) return;
  const tw::ChunkMeta cm = chunks[ch];
  chunk_row[(size_t)ch * tw::kMaxChunkDof + i] = i < cm.ndof ? row_out[chunk_dof[cm.dof_off + i]] : -1;
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
  {
    // per-launch temporaries come from the stream-ordered pool (cudaMalloc would synchronise the device, which
    // serialises against a peer's collective in multi-process runs); keep the pool's memory between launches
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
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
  {
    std::vector<int> dp(ps.ndof, 0);
    for (int p = 0; p < ps.npatch; p++)
      for (int i = ps.patch_dof_ptr[p]; i < ps.patch_dof_ptr[p + 1]; i++) dp[i] = p;
    if (!(e = upload(dp, &dof_patch)).empty()) return e;
  }
  {
    std::vector<ChunkAux> ax(ps.nchunk);
    std::memset(ax.data(), 0, ax.size() * sizeof(ChunkAux));
    for (int ch = 0; ch < ps.nchunk; ch++) {
      const ChunkMeta& cm = ps.chunks[ch];
      ChunkAux& x = ax[ch];
      for (int c = 0; c < kCH; c++) {
        x.dmin[c] = ps.cell_dmin[(size_t)ch * kCH + c];
        x.dmax[c] = ps.cell_dmax[(size_t)ch * kCH + c];
      }
      const int* ip = ps.chunk_inc_ptr.data() + cm.dof_off + ch;
      for (int i = 0; i <= cm.ndof; i++) x.iptr[i] = ip[i];
      for (int i = 0; i < ip[cm.ndof]; i++) x.inc[i] = ps.inc[(size_t)cm.inc_off + i];
      for (int i = 0; i < cm.ndof; i++) {
        x.orig[i] = ps.dof_orig[ps.chunk_dof[cm.dof_off + i]];
        unsigned hm = 0;
        for (int k = ip[i]; k < ip[i + 1]; k++) hm |= 1u << ((x.inc[k] & 63) / kRowHalf);
        for (int h = 0; h < kCH / kRowHalf; h++)
          if ((hm >> h) & 1u) x.act[h][x.nact[h]++] = (unsigned char)i;
      }
    }
    if (!(e = upload(ax, &aux)).empty()) return e;
    nchunk = ps.nchunk;
  }
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
  cudaFree(dof_patch);
  dof_patch = nullptr;
  cudaFree(aux);
  aux = nullptr;
  chunks = nullptr;
  geom = nullptr;
  dmin = dmax = chunk_dof = inc_ptr = patch_chunk_ptr = dof_orig = nullptr;
  inc = nullptr;
}

std::string gpu_symmetrize_cross(const DevicePatchSet& A, int i0, int i1, int j0, int j1, double* dst, const double* src, long long ld,
                                 cudaStream_t stream) {
  if (i1 <= i0 || j1 <= j0) return "";
  twk::symmetrize_cross_kernel<<<dim3((i1 - i0 + 31) / 32, (j1 - j0 + 31) / 32), dim3(32, 8), 0, stream>>>(i0, i1, j0, j1, A.dof_orig, dst, src, ld);
  CK(cudaGetLastError());
  note_launch();
  return "";
}

std::string gpu_lmat_tiles(const DevicePatchSet& A, const DevicePatchSet& B, const std::vector<Tile>& tiles,
                           const std::vector<int>& row_out, bool self, double* d_out, long long ld, cudaStream_t stream,
                           unsigned long long* h_stats, const int* d_col_map, bool symmetrize) {
  std::string e = gpu_init_constants();
  if (!e.empty()) return e;
  if (tiles.empty()) return "";
  Tile* d_tiles = nullptr;
  int* d_row_out = nullptr;
  int* d_counter = nullptr;
  unsigned long long* d_stats = nullptr;
  CK(cudaMallocAsync((void**)&d_tiles, tiles.size() * sizeof(Tile), stream));
  CK(cudaMallocAsync((void**)&d_row_out, std::max<size_t>(row_out.size(), 1) * sizeof(int), stream));
  CK(cudaMallocAsync((void**)&d_counter, sizeof(int), stream));
  CK(cudaMallocAsync((void**)&d_stats, 24 * sizeof(unsigned long long), stream));
  CK(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(d_row_out, row_out.data(), row_out.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  CK(cudaMemsetAsync(d_counter, 0, sizeof(int), stream));
  CK(cudaMemsetAsync(d_stats, 0, 24 * sizeof(unsigned long long), stream));
  CK(cudaMemsetAsync(d_stats + 7, 0xff, sizeof(unsigned long long), stream));
  int* d_chunk_row = nullptr;
  CK(cudaMallocAsync((void**)&d_chunk_row, (size_t)std::max(A.nchunk, 1) * kMaxChunkDof * sizeof(int), stream));
  twk::chunk_rows_kernel<<<std::max(A.nchunk, 1), kMaxChunkDof, 0, stream>>>(A.nchunk, A.chunks, A.chunk_dof, d_row_out, d_chunk_row);
  CK(cudaGetLastError());
  note_launch();
  twk::LmatArgs a;
  a.chunksA = A.chunks; a.chunksB = B.chunks;
  a.geomA = A.geom; a.geomB = B.geom;
  a.patch_chunk_ptrA = A.patch_chunk_ptr; a.patch_chunk_ptrB = B.patch_chunk_ptr;
  a.auxA = A.aux; a.auxB = B.aux;
  a.chunk_row = d_chunk_row;
  a.col_map = d_col_map;
  a.tiles = d_tiles;
  a.ntiles = (int)tiles.size();
  a.tile_counter = d_counter;
  a.out = d_out;
  a.ld = ld;
  a.scale = 1.0 / (4.0 * kPi);
  a.self = self ? 1 : 0;
  a.fast_lim = 64;
  a.debug_skip = 0;
#ifdef TW_TEST_HOOKS  // test / tuning build only (libthincurr_b200_test.so): force the rare drain paths, skip phases
  if (const char* e = std::getenv("THINCURR_B200_DRAIN_LIMIT")) a.fast_lim = std::atoi(e) >= 64 ? 64 : (std::atoi(e) >= 32 ? 32 : 0);
  a.debug_skip = std::getenv("THINCURR_B200_DEBUG_SKIP") ? std::atoi(std::getenv("THINCURR_B200_DEBUG_SKIP")) : 0;
#endif
  a.stats = d_stats;
  int dev = 0, nsm = 148;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  int grid = (int)std::min<size_t>(tiles.size(), (size_t)nsm);
  twk::lmat_tile_kernel<<<grid, twk::NT, sizeof(twk::Smem), stream>>>(a);
  CK(cudaGetLastError());
  note_launch();
  if (self && symmetrize) {
    // owned internal DOFs form one contiguous range (rows are numbered along it)
    int i0 = -1, i1 = -1;
    for (int i = 0; i < (int)row_out.size(); i++)
      if (row_out[i] >= 0) {
        if (i0 < 0) i0 = i;
        i1 = i + 1;
      }
    if (i0 >= 0 && i1 - i0 > 1) {
      const int nb = (i1 - i0 + 31) / 32;
      twk::symmetrize_kernel<<<dim3(nb, nb), dim3(32, 8), 0, stream>>>(i0, i1, A.dof_orig, A.dof_patch, d_out, ld);
      CK(cudaGetLastError());
      note_launch();
    }
  }
  if (h_stats) {
    unsigned long long hs[24];
    CK(cudaMemcpyAsync(hs, d_stats, 24 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    std::memcpy(h_stats, hs, 8 * sizeof(unsigned long long));
#ifdef TW_LMAT_PROF
    {
      static const char* nm[13] = {"tile_fetch", "wait+A0", "scatter+queue+bar", "eval", "drain_rest", "classify_loop", "bins+bar", "D1:contract", "D2:add_into_L", "-", "-", "-", "-"};
      std::fprintf(stderr, "[lmat prof, CTA 0, Mcycles]");
      for (int i = 0; i < 13; i++) std::fprintf(stderr, " %s=%.1f", nm[i], hs[8 + i] * 1e-6);
      std::fprintf(stderr, "\n");
    }
#endif
  }
  // stream-ordered frees keep the call asynchronous
  CK(cudaFreeAsync(d_tiles, stream));
  CK(cudaFreeAsync(d_row_out, stream));
  CK(cudaFreeAsync(d_chunk_row, stream));
  CK(cudaFreeAsync(d_counter, stream));
  CK(cudaFreeAsync(d_stats, stream));
  return "";
}

}  // namespace tw
