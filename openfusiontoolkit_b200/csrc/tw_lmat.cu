// tw_lmat.cu -- dense element<->element inductance build on sm_100a (FP64, no tensor cores).
//
// Replaces the O(nc^2) OpenMP loop nest of tw_compute_LmatDirect (src/physics/thin_wall.F90:
// 1008-1126) with an owner-computes tiling.  One persistent CTA per SM pulls output tiles
// (row patch x column patch) from a cost-sorted queue.  One PASS = one pair of 64-cell chunks of the
// two patches (4096 cell pairs):
//   0. staging: a chunk is three bulk async copies (cp.async.bulk ... mbarrier::complete_tx) into a
//      ChunkState slot -- 22 SoA geometry rows, the index record (per-cell min/max DOF, reference DOF ids,
//      CSR incidences) and this launch's output rows; the next column chunk is prefetched during the pass.
//      All local coordinates of a sweep are relative to the centre of the ROW chunk, so its FP32 copies and
//      its quadrature-point tables (rules of 6, 7 and 12 points) are built once per row chunk;
//   1. the column chunk's FP32 copies and point tables (one barrier);
//   2. ROW SWEEP, no barrier inside: a warp takes one row cell at a time from a counter, classifies its 64
//      pairs -- the quadrature order of thin_wall.F90:1044-1059 screened in FP32 with a rigorous error band
//      (order 4 from the cells' bounding spheres when they are far enough apart; bit-exact FP64 evaluation
//      when a threshold is within the band; nothing at all when the chunks' bounding spheres already
//      guarantee order 4 for every pair) -- compacts the pairs of the three tabled rules with warp votes
//      into its own pending lists and evaluates every full batch of 32 at once from the shared-memory
//      point tables ((x,y,z,|x|^2), d^2 = |xi|^2+|xj|^2-2 xi.xj, MUFU.RSQ64H seed + third-order
//      correction = 10 FP64-pipe instructions per 1/r).  Warps drift apart, so the FP32/integer work of
//      the classifying warps overlaps the FP64 work of the evaluating ones;
//   3. what is left -- the warps' partial batches, pairs of the larger far rules (evaluated from the
//      vertices) and near pairs (thin_wall.F90:1061-1068: one half-warp per pair, lanes over the points
//      of the analytic potential) -- is evaluated from block-wide lists through a dynamic queue;
//   4. contraction onto the vertex/hole DOFs, fused with the update of L: a warp per row DOF sums its
//      cells' rows of T against the column side and adds the result into its row of L with lanes along
//      the row (old values loaded up front).
// Every L entry is owned by exactly one CTA and updated by plain loads and stores between barriers: no
// atomics on the matrix, deterministic values.  When the rows of both patches of a tile are in the output
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
constexpr int kTabMaxN = 25;       // largest rule far_tab is instantiated for (probes)
constexpr int kTabClsMax = 2;      // far classes served from the shared-memory point tables by the tile kernel (rules 4, 5, 6)
constexpr int kTabPts = 25;        // 6 + 7 + 12 points
constexpr int kPend = 64;          // pending pairs per (warp, tabled class): < 32 left over + <= 32 of a new unit
constexpr int kLeft = 512;         // leftover pairs per tabled class and pass (< 32 per warp)
constexpr int kOList = CI * kCH + (NCLS - 3) * 32;  // pairs of the other classes, bins padded to batches

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
  int fast_lim;                       // column DOFs per chunk handled from registers in the contraction (64; 32 or 0 in tests of the generic path)
  int debug_skip;                     // test build only: bit0 skip near-field evaluation, bit1 skip far-field evaluation,
                                      // bit2 skip the contraction
  unsigned long long* stats;          // [0] far pairs, [1] near T evaluations, [2] 1/r evaluations, [3] phipot evals
  // streamed build (one launch, matrix in the reference layout, tiles ordered band by band; nbands = 0 otherwise):
  int nbands;
  const int* tile_band;               // [ntiles] band of a tile (the queue order is free: latest start time first)
  const int* band_ntiles;             // [nbands] tiles per band
  const int* band_ref;                // [nbands][2] first / one-past-last reference DOF id of the band's rows
  const int* ref_patch;               // [N] patch of a reference DOF id
  int N;
  int* band_state;                    // [3][nbands]: tiles finished, next mirror task, mirror tasks finished
  volatile int* host_flags;           // [nbands] (mapped host memory) set when the band's part of the matrix is final
};
constexpr int kMirKC = 256;           // 32-column blocks per mirror task

// one staged chunk (row or column side): SoA geometry record + index record + output rows, each
// filled by one bulk async copy
struct alignas(16) ChunkState {
  double g[kGeomL * kCH];             // rows 0-8 vertices, 9 area, 10-18 qbasis, 19-21 unit normal (phipot's)
  tw::ChunkAux x;
  int row[tw::kMaxChunkDof];          // output rows (or -1)
  double cx, cy, cz, rad, emax;
  int ncell, ndof;
  unsigned long long bar;             // mbarrier of the bulk copies
};
static_assert(offsetof(ChunkState, x) % 16 == 0 && offsetof(ChunkState, row) % 16 == 0, "bulk copy alignment");

// per-pass work lists
struct alignas(16) PassBuf {
  unsigned short olist[kOList];       // pairs (c1<<6|c2) of the classes > kTabClsMax sorted by class, bins padded with 0xFFFF
  unsigned short left[kTabClsMax + 1][kLeft];       // partial batches of the tabled classes handed over by the warps
  unsigned short pend[NW][kTabClsMax + 1][kPend];   // per warp: pairs of the tabled classes waiting for a full batch
  unsigned char iqmap[CI * kCH];      // per pair: 0, or for a pair of another class iquad | need-role-1 << 5 | need-role-2 << 6
  int nleft[kTabClsMax + 1];
  int cnt[NCLS], pos[NCLS], off[NCLS + 1];
  int qcls[NCLS], qnb[NCLS];          // queue items: class (| 16: whole-warp batches from on-demand tables), batches
  int gq0[4], gq1[4], gnb[4], gcls[4], ngrp;  // groups: item range, batch count, tabled class (or -1)
  int rowhead, lhead, qhead, both_count;
};

struct Smem {
  ChunkState I;                       // row chunk
  ChunkState J[2];                    // column chunks: the one in use and the next one (prefetched)
  PassBuf pb;
  alignas(16) double T[CI * TS];
  alignas(16) double2 tabI[kTabPts * 2 * CI];   // [(p*2+h)*CI + c1]: h=0 (-2x,-2y), h=1 (-2z,|x|^2); rules 4,5,6 at points 0,6,13
  union {
    double2 tabJ[kTabPts * 2 * kCH];  //          [(p*2+h)*kCH + c2]: h=0 (x,y),    h=1 (z,|x|^2)
    double P[NW][3 * kCH];            // contraction scratch of the warps (the column tables are dead by then)
  } u;
  float vfI[9 * CI], vfJ[9 * kCH];    // vertices relative to the row chunk's centre, FP32 (order screening)
  float flI[CI], flJ[kCH];            // 2 * area
  float4 cenI[CI], cenJ[kCH];         // centroid (same frame) and the radius covering the vertices
  int item[5];                        // [0] tile (-1: exit, -2: mirror task [1] of band [2]), [3], [4]: see fetch_banded
};
static_assert(sizeof(Smem) <= 232448, "shared memory of one CTA");
static_assert(offsetof(Smem, tabI) == offsetof(Smem, T) + sizeof(double) * CI * TS && offsetof(Smem, u) == offsetof(Smem, tabI) + sizeof(double2) * kTabPts * 2 * CI &&
                  sizeof(double) * CI * TS + 2 * sizeof(double2) * kTabPts * 2 * CI >= NW * 32 * 33 * sizeof(double),
              "T, tabI and tabJ form the contiguous scratch of the mirror tasks (one 32 x 33 block per warp)");
template <int N> struct ShowSize;
#ifdef TW_SHOW_SMEM
ShowSize<sizeof(Smem)> show_smem_size;
#endif
#define TW_MARK(S, who, t, i)

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

// ---- far field of the larger rules: a group of G lanes per pair -------------------------------------------------
// The rules of 15..25 points are rare per pass (a few dozen pairs at the edge of the near field), so a pass cannot fill
// warps with 32 pairs of one of them, and whole-warp batches would leave most warps idle.  G lanes share a pair: lane g
// holds the column points q = g, g+G, ... in registers (frame centred on the row cell, (x,y,z,|x|^2)), the row points are
// computed once per group (lane g takes p = g, g+G, ...) and handed round with shuffles, and the 1/r evaluation is the
// same 10-instruction form as the table path.  Fixed summation order: deterministic.  All 32 lanes must call (shuffles).
template <int N, int G>  // @region far_group
__device__ __forceinline__ double far_group(const double* __restrict__ gI, int c1, const double* __restrict__ gJ, int c2, int iquad, int lane) {
  constexpr int QG = (N + G - 1) / G;
  const int g = lane & (G - 1), base = lane & ~(G - 1);
  const double* bp = g_qpts + 3 * c_qoff[iquad];  // lane-dependent point index: global copy of the tables
  const double* bwg = g_qwts + c_qoff[iquad];
  const double* bwc = c_qwts + c_qoff[iquad];     // uniform index: constant bank
  double Pi[9], Pj[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    Pi[k] = gI[k * kCH + c1];
    Pj[k] = gJ[k * kCH + c2];
  }
  const double ox = (Pi[0] + Pi[3] + Pi[6]) * (1.0 / 3.0), oy = (Pi[1] + Pi[4] + Pi[7]) * (1.0 / 3.0), oz = (Pi[2] + Pi[5] + Pi[8]) * (1.0 / 3.0);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    Pi[3 * k] -= ox;
    Pi[3 * k + 1] -= oy;
    Pi[3 * k + 2] -= oz;
    Pj[3 * k] -= ox;
    Pj[3 * k + 1] -= oy;
    Pj[3 * k + 2] -= oz;
  }
  double xj[QG], yj[QG], zj[QG], sj[QG], acc[QG];
#pragma unroll
  for (int m = 0; m < QG; m++) {
    const int q = min(m * G + g, N - 1);  // the tail re-reads the last point (weight masked below)
    const double b0 = bp[3 * q], b1 = bp[3 * q + 1], b2 = bp[3 * q + 2];
    xj[m] = b0 * Pj[0] + b1 * Pj[3] + b2 * Pj[6];
    yj[m] = b0 * Pj[1] + b1 * Pj[4] + b2 * Pj[7];
    zj[m] = b0 * Pj[2] + b1 * Pj[5] + b2 * Pj[8];
    sj[m] = fma(zj[m], zj[m], fma(yj[m], yj[m], xj[m] * xj[m]));
    acc[m] = 0.0;
  }
#pragma unroll 1
  for (int p0 = 0; p0 < N; p0 += G) {
    // this lane's row point p0 + g as (-2x, -2y, -2z, |x|^2)
    const int p = min(p0 + g, N - 1);
    const double a0 = bp[3 * p], a1 = bp[3 * p + 1], a2 = bp[3 * p + 2];
    const double xi = a0 * Pi[0] + a1 * Pi[3] + a2 * Pi[6], yi = a0 * Pi[1] + a1 * Pi[4] + a2 * Pi[7], zi = a0 * Pi[2] + a1 * Pi[5] + a2 * Pi[8];
    const double si = fma(zi, zi, fma(yi, yi, xi * xi));
    const double mx = -2.0 * xi, my = -2.0 * yi, mz = -2.0 * zi;
#pragma unroll
    for (int kk = 0; kk < G; kk++) {
      if (p0 + kk < N) {  // uniform
        const double ax = __shfl_sync(0xffffffffu, mx, base + kk), ay = __shfl_sync(0xffffffffu, my, base + kk),
                     az = __shfl_sync(0xffffffffu, mz, base + kk), as = __shfl_sync(0xffffffffu, si, base + kk);
        const double wp = bwc[p0 + kk];
        double d2[QG], y0[QG], e[QG], h[QG];
#pragma unroll
        for (int m = 0; m < QG; m++) d2[m] = fma(ax, xj[m], fma(ay, yj[m], fma(az, zj[m], as + sj[m])));
#pragma unroll
        for (int m = 0; m < QG; m++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[m]) : "d"(d2[m]));
#pragma unroll
        for (int m = 0; m < QG; m++) h[m] = d2[m] * y0[m];
#pragma unroll
        for (int m = 0; m < QG; m++) e[m] = fma(-h[m], y0[m], 1.0);
#pragma unroll
        for (int m = 0; m < QG; m++) h[m] = fma(0.375, e[m], 0.5);
#pragma unroll
        for (int m = 0; m < QG; m++) e[m] = e[m] * y0[m];
#pragma unroll
        for (int m = 0; m < QG; m++) y0[m] = fma(e[m], h[m], y0[m]);
#pragma unroll
        for (int m = 0; m < QG; m++) acc[m] = fma(wp, y0[m], acc[m]);
      }
    }
  }
  double total = 0.0;
#pragma unroll
  for (int m = 0; m < QG; m++)
    if (m * G + g < N) total = fma(bwg[m * G + g], acc[m], total);
#pragma unroll
  for (int o = 1; o < G; o <<= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  return total * gI[9 * kCH + c1] * gJ[9 * kCH + c2];
}

// pairs per warp batch of a class of the block-wide lists: larger far rules in lane groups, near pairs in half-warps
// (near pairs: whole warps while a pass holds only a few of them -- the finish time of the phase matters more than the
// idle lanes of the short rules)
__device__ __forceinline__ int batch_pairs(int cls, int near_ppb) { return cls >= 7 ? near_ppb : (cls == 6 ? 4 : 8); }
__device__ __forceinline__ int near_batch_pairs(const int* cnt);
// a larger far rule with at least kTabMin pairs in the pass is evaluated from point tables built on demand (whole-warp
// batches); such pairs come in bulk in the passes between nearby chunks
constexpr int kTabMin = 64;
__device__ __forceinline__ bool big_rule_tabled(int cls, const int* cnt) { return cls > kTabClsMax && cls < 7 && cnt[cls] >= kTabMin; }
__device__ __forceinline__ int list_batch_pairs(int cls, const int* cnt, int near_ppb) {
  return big_rule_tabled(cls, cnt) ? 32 : batch_pairs(cls, near_ppb);
}
__device__ __forceinline__ int near_batch_pairs(const int* cnt) {
  int n = 0;
#pragma unroll
  for (int c = 7; c < NCLS; c++) n += cnt[c];
  return n < 96 ? 1 : 2;
}

__device__ __noinline__ double far_group_dispatch(const double* gI, int c1, const double* gJ, int c2, int cls, int lane) {  // @region far_group_dispatch
  switch (cls) {
    case 3: return far_group<15, 4>(gI, c1, gJ, c2, 7, lane);
    case 4: return far_group<16, 4>(gI, c1, gJ, c2, 8, lane);
    case 5: return far_group<19, 4>(gI, c1, gJ, c2, 9, lane);
    default: return far_group<25, 8>(gI, c1, gJ, c2, 10, lane);
  }
}

// ---- far field from the shared-memory point tables ----------------------------------------------
// 10 FP64-pipe instructions per 1/r: 1 add + 3 fma (d^2), 5 (rsqrt correction), 1 fma (weighted sum).
// The QB evaluations of one row point are advanced stage by stage so that QB independent
// dependency chains are in flight (DFMA latency is 8 cycles, the pipe takes one warp every 2).
// one block of QB column points (q0 .. q0+QB-1, all valid) against all N row points; `total` carries the weighted sum
template <int N, int OFF, int QB>  // @region far_tab
__device__ __forceinline__ double far_tab_block(const double2* __restrict__ tabI, const double2* __restrict__ tabJ, int c1, int c2, int q0, double total) {
  const double* bw = c_qwts + OFF;  // OFF = TCQ_OFF[iquad]: weights become constant-bank operands
  double xj[QB], yj[QB], zj[QB], sj[QB], acc[QB];
#pragma unroll
  for (int q = 0; q < QB; q++) {
    const double2 u = tabJ[((q0 + q) * 2) * kCH + c2], v = tabJ[((q0 + q) * 2 + 1) * kCH + c2];
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
  for (int q = 0; q < QB; q++) total = fma(bw[q0 + q], acc[q], total);
  return total;
}
// N column points in blocks of QB held in registers, the remainder (7 = 4 + 3, 19 = 4 x 4 + 3) as a block of its own size
template <int N, int OFF>
__device__ __forceinline__ double far_tab(const double2* __restrict__ tabI, const double2* __restrict__ tabJ, int c1, int c2) {
  // c1: row cell within the pass (stride CI), c2: column cell (stride kCH)
  constexpr int QB = (N == 6) ? 3 : ((N == 15 || N == 25) ? 5 : 4);
  constexpr int NFULL = N / QB, REM = N - NFULL * QB;
  double total = 0.0;
#pragma unroll 1
  for (int q0 = 0; q0 < NFULL * QB; q0 += QB) total = far_tab_block<N, OFF, QB>(tabI, tabJ, c1, c2, q0, total);
  if (REM > 0) total = far_tab_block<N, OFF, (REM > 0 ? REM : 1)>(tabI, tabJ, c1, c2, NFULL * QB, total);
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
  C.emax = cm.emax;
}

// ---- preparation of a chunk for the sweep: FP32 copies (order screening) and point tables of the tabled rules, all
// relative to the centre (ox,oy,oz) of the ROW chunk.  row side: once per row chunk; column side: once per pass.
__device__ __forceinline__ void prep_chunk(const ChunkState& C, float* __restrict__ vf, float* __restrict__ fl, float4* __restrict__ cen,  // @region prep_chunk
                                           double2* __restrict__ tab, bool neg, double ox, double oy, double oz, int tid) {
  for (int i = tid; i < 9 * kCH; i += NT) {
    const int dd = (i / kCH) % 3;
    vf[i] = (float)(C.g[i] - (dd == 0 ? ox : (dd == 1 ? oy : oz)));
  }
  if (tid >= NT - kCH) {  // bounding sphere of every cell (order-4 prefilter of the classification) and 2 * area
    const int c = tid - (NT - kCH);
    double P[9];
#pragma unroll
    for (int k = 0; k < 9; k++) P[k] = C.g[k * kCH + c];
    const double mx = (P[0] + P[3] + P[6]) * (1.0 / 3.0), my = (P[1] + P[4] + P[7]) * (1.0 / 3.0), mz = (P[2] + P[5] + P[8]) * (1.0 / 3.0);
    // radius over the vertices; the dl_max floor sqrt(2 area) <= 1.62 x this radius needs no separate term: the test
    // of the prefilter implies D - R > 79 R
    double r2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double dx = P[3 * k] - mx, dy = P[3 * k + 1] - my, dz = P[3 * k + 2] - mz;
      r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
    }
    cen[c] = make_float4((float)(mx - ox), (float)(my - oy), (float)(mz - oz), (float)(sqrt(r2) * 1.00001));
    fl[c] = (float)(2.0 * C.g[9 * kCH + c]);
  }
  build_table(tab, kCH, C.g, 0, C.ncell, 4, 6, ox, oy, oz, neg, tid, NT);
  build_table(tab + 6 * 2 * kCH, kCH, C.g, 0, C.ncell, 5, 7, ox, oy, oz, neg, tid, NT);
  build_table(tab + 13 * 2 * kCH, kCH, C.g, 0, C.ncell, 6, 12, ox, oy, oz, neg, tid, NT);
}
static_assert(CI == kCH, "row and column tables share the stride");

// one batch of a tabled class: lane's pair e (c1<<6|c2, 0xFFFF = none) from the point tables
template <int C>
__device__ __forceinline__ void tab_eval(Smem& S, const ChunkState& I, const ChunkState& J, unsigned e) {  // @region tab_eval
  constexpr int N = C == 0 ? 6 : (C == 1 ? 7 : 12), OFF = C == 0 ? 7 : (C == 1 ? 13 : 20), PT = C == 0 ? 0 : (C == 1 ? 6 : 13);
  if (e != 0xFFFFu) {
    const int c1 = e >> 6, c2 = e & 63;
    S.T[c1 * TS + c2] = far_tab<N, OFF>(S.tabI + PT * 2 * CI, S.u.tabJ + PT * 2 * kCH, c1, c2) * I.g[9 * kCH + c1] * J.g[9 * kCH + c2];
  }
}

// append the lanes' pairs of class C (flag `mine`, pair id `e`) to the warp's pending list and evaluate full batches
template <int C>
__device__ __forceinline__ void tab_push(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, unsigned short* __restrict__ pend, int& npd,
                                         bool mine, unsigned e, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  if (m == 0u) return;
  if (mine) pend[npd + __popc(m & ((1u << lane) - 1u))] = (unsigned short)e;
  npd += __popc(m);
  __syncwarp();
  if (npd >= 32) {
    npd -= 32;
    const unsigned eb = pend[npd + lane];
    __syncwarp();  // the slots may be overwritten by the next append
    if (!(A.debug_skip & 2)) tab_eval<C>(S, I, J, eb);
  }
}

// ---- row sweep: classification of the pairs of one row cell at a time, tabled rules evaluated at once --------------
__device__ __forceinline__ void sweep_rows(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, int flags, int tid,  // @region sweep_rows
                                           unsigned long long& st_far, unsigned long long& st_eval) {
  PassBuf& pb = S.pb;
  const int lane = tid & 31, warp = tid >> 5;
  const int ncI = I.ncell, ncJ = J.ncell;
  const bool diag = flags & 1, want2 = (flags & 4) && A.self;
  // block-uniform geometry of the pass
  const double hx = J.cx - I.cx, hy = J.cy - I.cy, hz = J.cz - I.cz;
  const double D = sqrt(hx * hx + hy * hy + hz * hz);
  // |v| <= X for every FP32 coordinate of the pass; delta = bound of the error of a vertex difference (two operands, each
  // rounded to FP32: 2^-24 relative, 1% slack)
  const float delta = (float)((D + fmax(I.rad, J.rad)) * 1.21e-7);
  const double gap = D - I.rad - J.rad;
  // every pair of the pass has order 4: dl_min >= gap, dl_max <= dl_min + emax_I + emax_J (vertices of a triangle are within
  // one edge of each other; the floor sqrt(2 area) is below one edge), so rho >= gap / (gap + s) > c_thr[0]
  const double s_e = I.emax + J.emax;
  const bool uniform4 = gap > 0.0 && gap * (1.0 - c_thr[0]) > s_e * c_thr[0] * (1.0 + 1.0e-9);
  // the per-pair order-4 prefilter pays only where most pairs pass: chunks at least ~25 cell sizes apart
  const bool far_pass = gap > 25.0 * (double)(S.cenI[0].w + S.cenJ[0].w);
  unsigned short* pend0 = pb.pend[warp][0];
  unsigned short* pend1 = pb.pend[warp][1];
  unsigned short* pend2 = pb.pend[warp][2];
  int np0 = 0, np1 = 0, np2 = 0;
#pragma unroll 1
  for (;;) {
    // unit = half a row: 32 pairs (c1, c2 = lane + 32 h), one per lane (small units keep the warps' finish times close)
    int u = 0;
    if (lane == 0) u = atomicAdd(&pb.rowhead, 1);
    u = __shfl_sync(0xffffffffu, u, 0);
    const int c1 = u >> 1;
    if (c1 >= ncI) break;
    const int c2 = lane + 32 * (u & 1);
    int cls = -1;
    unsigned code = 0;
    if (c2 < ncJ) {
      bool n1, n2 = false;
      if (A.self) {
        n1 = I.x.dmin[c1] <= J.x.dmax[c2];
        n2 = want2 && !diag && (I.x.dmax[c1] > J.x.dmin[c2]);
      } else {
        n1 = true;
      }
      if (n1 || n2) {
        int iq = 4;
        if (!uniform4) {
          float pi_[9], pj_[9];
#pragma unroll
          for (int k = 0; k < 9; k++) {
            pi_[k] = S.vfI[k * CI + c1];
            pj_[k] = S.vfJ[k * kCH + c2];
          }
          iq = -2;
          if (far_pass) {
            // order-4 prefilter.  u = unit vector between the centroids, a_k / b_k = projections of the vertices on u,
            // R = sum of the cells' vertex radii (>= the part of any vertex difference perpendicular to u):
            //   dl_min >= Lb = min b - max a,   dl_max <= Ub = (max b - min a) + R^2 / (2 Lb)
            // (the floor sqrt(2 area) <= 1.62 R is below Ub once D > 10 R), so Lb/Ub > 0.9752 > c_thr[0] => iquad = 4.
            const float4 ci = S.cenI[c1], cj = S.cenJ[c2];
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
            iq = iquad_screen(pi_, pj_, fmaxf(S.flI[c1], S.flJ[c2]), delta);
            if (iq < 0) iq = iquad_exact_cells(I.g, c1, J.g, c2);
          }
        }
        cls = cls_of(iq);
        if (cls > kTabClsMax) code = (unsigned)iq | (n1 ? 32u : 0u) | (n2 ? 64u : 0u);
        if (cls < 7) {  // statistics: far pairs and 1/r evaluations
          st_far++;
          st_eval += (unsigned long long)(c_cls_np[cls] * c_cls_np[cls]);
        }
      }
    }
    pb.iqmap[c1 * kCH + c2] = (unsigned char)code;
    // tabled rules: compact with warp votes, evaluate every full batch
    const unsigned e = (unsigned)(c1 * kCH + c2);
    tab_push<0>(S, A, I, J, pend0, np0, cls == 0, e, lane);
    tab_push<1>(S, A, I, J, pend1, np1, cls == 1, e, lane);
    tab_push<2>(S, A, I, J, pend2, np2, cls == 2, e, lane);
    // other classes: counts for the block-wide bins
    const unsigned mo = __ballot_sync(0xffffffffu, code != 0u);
    if (mo) {
      int mycount = 0;
#pragma unroll
      for (int c = kTabClsMax + 1; c < NCLS; c++) {
        const int k = __popc(__ballot_sync(0xffffffffu, cls == c));
        if (lane == c) mycount = k;
      }
      if (mycount) atomicAdd(&pb.cnt[lane], mycount);
      const int nboth = __popc(__ballot_sync(0xffffffffu, (code & 96u) == 96u));
      if (lane == 0 && nboth) atomicAdd(&pb.both_count, nboth);
    }
  }
  // hand the partial batches over
  {
    const int npd[3] = {np0, np1, np2};
    unsigned short* const pd[3] = {pend0, pend1, pend2};
#pragma unroll
    for (int c = 0; c <= kTabClsMax; c++) {
      if (npd[c] == 0) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&pb.nleft[c], npd[c]);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (lane < npd[c]) pb.left[c][base + lane] = pd[c][lane];
    }
  }
}

// ---- block-wide bins of the other classes (larger far rules, near pairs) from the per-pair codes ----------------------
// pb.cnt is complete (barrier before).  A thread takes the codes of 8 consecutive pairs of a row; packed histograms +
// one warp scan give the lane offsets, one shared atomic per (warp, class) the warp's share of the bin.  Bins: near
// classes (largest rules) first, padded to whole batches.  The last warp builds the queue.  A barrier must follow.
__device__ __forceinline__ void bin_others(Smem& S, const ChunkState& I, int tid) {  // @region bin_others
  PassBuf& pb = S.pb;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int TPR = NT / CI, NIT = kCH / TPR;
  static_assert(NIT == 8 && TPR * CI == NT, "8 pair codes per thread");
  const int c1 = tid / TPR, c2b = NIT * (tid % TPR);
  unsigned long long codes8 = 0;
  if (c1 < I.ncell) codes8 = *reinterpret_cast<const unsigned long long*>(pb.iqmap + c1 * kCH + c2b);
  unsigned mycls = 0;            // 4 bits per pair: class + 1, 0 = not in a bin
  unsigned long long hist = 0;   // 4 bits per class
#pragma unroll
  for (int m = 0; m < NIT; m++) {
    const unsigned cd = (unsigned)(codes8 >> (8 * m)) & 255u;
    if (cd) {
      const int cls = cls_of((int)(cd & 31u));
      mycls |= (unsigned)(cls + 1) << (4 * m);
      hist += 1ull << (4 * cls);
    }
  }
  unsigned long long h0 = 0, h1 = 0;  // 10-bit fields, 6 classes per word
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
  int start = 0;  // lane c < NCLS: first list slot of this warp's pairs of class c
  if (lane < NCLS) {
    const int tot = (int)(((lane < 6 ? w0 : w1) >> (10 * (lane % 6))) & 1023ull);
    int wbase = 0;
    if (tot) wbase = atomicAdd(&pb.pos[lane], tot);
    int o = 0;
    const int nppb = near_batch_pairs(pb.cnt);
    for (int c = NCLS - 1; c > lane; c--) {
      const int n = pb.cnt[c], bp_ = list_batch_pairs(c, pb.cnt, nppb);
      o += (n + bp_ - 1) / bp_ * bp_;
    }
    start = o + wbase;
    if (warp == 0) {
      const int n = pb.cnt[lane], bp_ = list_batch_pairs(lane, pb.cnt, nppb);
      pb.off[lane] = o;
      const int padded = (n + bp_ - 1) / bp_ * bp_;
      for (int i = n; i < padded; i++) pb.olist[o + i] = 0xFFFFu;  // padding of the last batch
    }
  }
  if (warp == NW - 1) {
    // queue.  Group 0: near classes (largest rules first) and the small bins of the larger far rules (lane groups), then
    // the first tabled rule; every further tabled rule is a group of its own (a group is one barrier interval: its point
    // tables occupy the table memory).  Items in descending class order.
    const int cl = lane < NCLS ? lane : 0;
    const int n = (lane < NCLS && lane > kTabClsMax) ? pb.cnt[lane] : 0;
    const int nppb = near_batch_pairs(pb.cnt);
    const bool tabd = n > 0 && big_rule_tabled(cl, pb.cnt);
    const int bp_ = list_batch_pairs(cl, pb.cnt, nppb);
    const int nbat = (n + bp_ - 1) / bp_;
    const unsigned m_c0 = __ballot_sync(0xffffffffu, n > 0 && !tabd), m_tab = __ballot_sync(0xffffffffu, tabd);
    const int nq0 = __popc(m_c0), ntab = __popc(m_tab);
    if (n > 0 && !tabd) {
      const int q = __popc(m_c0 >> (lane + 1));
      pb.qcls[q] = lane;
      pb.qnb[q] = nbat;
    }
    if (tabd) {
      const int q = nq0 + __popc(m_tab >> (lane + 1));
      pb.qcls[q] = lane | 16;  // bit 4: from the tables
      pb.qnb[q] = nbat;
    }
    int nb0 = (n > 0 && !tabd) ? nbat : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nb0 += __shfl_xor_sync(0xffffffffu, nb0, o);
    if (tabd) {  // group of this rule: the first tabled rule shares group 0
      const int k = __popc(m_tab >> (lane + 1));  // index among the tabled rules (descending)
      const int g = k;                            // group index
      pb.gq0[g] = k == 0 ? 0 : nq0 + k;
      pb.gq1[g] = nq0 + k + 1;
      pb.gnb[g] = nbat + (k == 0 ? nb0 : 0);
      pb.gcls[g] = lane;
    }
    if (lane == 0) {
      if (ntab == 0) {
        pb.gq0[0] = 0;
        pb.gq1[0] = nq0;
        pb.gnb[0] = nb0;
        pb.gcls[0] = -1;
      }
      pb.ngrp = ntab > 0 ? ntab : 1;
    }
  }
  unsigned long long used = 0;  // 4 bits per class: pairs of this thread already placed
#pragma unroll 1
  for (int m = 0; m < NIT; m++) {
    const int cls = (int)((mycls >> (4 * m)) & 15u) - 1;
    const int st = __shfl_sync(0xffffffffu, start, cls < 0 ? 0 : cls);
    if (cls >= 0) {
      const int excl = (int)(((cls < 6 ? s0 : s1) >> (10 * (cls % 6))) & 1023ull);
      const int mine = (int)((used >> (4 * cls)) & 15ull);
      used += 1ull << (4 * cls);
      pb.olist[st + excl + mine] = (unsigned short)(c1 * kCH + c2b + m);
    }
  }
}

// ---- evaluation from the block-wide lists ------------------------------------------------------------------------
// one batch of the other classes: 32 far pairs of one rule (from the vertices) or 2 near pairs.
// List entries are (c1<<6 | c2).
__device__ __forceinline__ void run_batch_c0(const ChunkState& I, const ChunkState& J, const PassBuf& pb, double* __restrict__ T,  // @region run_batch_c0
                                             int cls, int first, int lane, bool role2_pass, int nppb, unsigned long long& st_near,
                                             unsigned long long& st_phi) {
  if (cls < 7) {
    const int G = cls == 6 ? 8 : 4;  // lanes per pair
    const unsigned e = pb.olist[first + lane / G];
    const bool ok = e != 0xFFFFu;
    const int c1 = ok ? (int)(e >> 6) : 0, c2 = ok ? (int)(e & 63) : 0;
    const double v = far_group_dispatch(I.g, c1, J.g, c2, cls, lane);
    if (ok && (lane & (G - 1)) == 0) T[c1 * TS + c2] = v;
  } else {
    const int nl = nppb == 1 ? 32 : 16;                 // lanes per pair
    const int hw = nppb == 1 ? 0 : lane >> 4, hl = lane & (nl - 1);
    const unsigned e = pb.olist[first + hw];
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
      const unsigned mask = nppb == 1 ? 0xffffffffu : 0xFFFFu << (16 * hw);
      double v;
      if (do2) v = near_pair(J.g, J.g + 19 * kCH, c2, I.g, c1, iq, hl, nl, mask);
      else v = near_pair(I.g, I.g + 19 * kCH, c1, J.g, c2, iq, hl, nl, mask);
      if (hl == 0) {
        T[c1 * TS + c2] = v;
        st_near++;
        st_phi += c_qnp[iq];
      }
    }
  }
}

// partial batches of the tabled classes handed over by the warps (complete after the sweep's barrier)
__device__ __forceinline__ void eval_leftovers(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, int tid) {  // @region eval_leftovers
  PassBuf& pb = S.pb;
  const int lane = tid & 31;
  const int n0 = pb.nleft[0], n1 = pb.nleft[1], n2 = pb.nleft[2];
  const int b0 = (n0 + 31) >> 5, b1 = (n1 + 31) >> 5, b2 = (n2 + 31) >> 5;
  const int total = b0 + b1 + b2;
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(&pb.lhead, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= total) break;
    if (A.debug_skip & 2) continue;
    if (b < b0) {
      const int i = b * 32 + lane;
      tab_eval<0>(S, I, J, i < n0 ? pb.left[0][i] : 0xFFFFu);
    } else if (b < b0 + b1) {
      const int i = (b - b0) * 32 + lane;
      tab_eval<1>(S, I, J, i < n1 ? pb.left[1][i] : 0xFFFFu);
    } else {
      const int i = (b - b0 - b1) * 32 + lane;
      tab_eval<2>(S, I, J, i < n2 ? pb.left[2][i] : 0xFFFFu);
    }
  }
}

// other classes from the sorted lists through the dynamic queue, group g
__device__ __forceinline__ void eval_others(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, int tid, int g,  // @region eval_others
                                            unsigned long long& st_near, unsigned long long& st_phi) {
  PassBuf& pb = S.pb;
  const int lane = tid & 31;
  const int q0 = pb.gq0[g], total = pb.gnb[g];
  const int nppb = near_batch_pairs(pb.cnt);
  int bnext = 0;
  if (lane == 0) bnext = atomicAdd(&pb.qhead, 1);
  for (;;) {
    const int b = __shfl_sync(0xffffffffu, bnext, 0);
    if (b >= total) break;
    if (lane == 0) bnext = atomicAdd(&pb.qhead, 1);  // the next batch index is fetched while this one is evaluated
    int k = q0, lb = b;
    while (lb >= pb.qnb[k]) {
      lb -= pb.qnb[k];
      k++;
    }
    const int qc = pb.qcls[k], cls = qc & 15;
    if (A.debug_skip && ((cls >= 7) ? (A.debug_skip & 1) : (A.debug_skip & 2))) continue;
    if (qc & 16) {
      const unsigned e = pb.olist[pb.off[cls] + lb * 32 + lane];
      if (e != 0xFFFFu) {
        const int c1 = e >> 6, c2 = e & 63;
        S.T[c1 * TS + c2] = far_tab_dispatch(S.tabI, S.u.tabJ, c1, c2, cls) * I.g[9 * kCH + c1] * J.g[9 * kCH + c2];
      }
    } else {
      run_batch_c0(I, J, pb, S.T, cls, pb.off[cls] + lb * batch_pairs(cls, nppb), lane, false, nppb, st_near, st_phi);
    }
  }
}

// second role: re-evaluate the near pairs that need both roles (all other T values stay)
__device__ __forceinline__ void eval_role2(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, int tid,  // @region C_role2
                                           unsigned long long& st_near, unsigned long long& st_phi) {
  PassBuf& pb = S.pb;
  const int lane = tid & 31;
  const int nppb = near_batch_pairs(pb.cnt);
  int nearb = 0, first_cls_off[5], first_cls_nb[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    first_cls_off[k] = pb.off[11 - k];
    first_cls_nb[k] = (pb.cnt[11 - k] + nppb - 1) / nppb;
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
    run_batch_c0(I, J, pb, S.T, 11 - k, first_cls_off[k] + lb * nppb, lane, true, nppb, st_near, st_phi);
  }
}

// ---- contraction of T onto the DOFs, fused with the update of L -------------------------------------------------
// A warp takes row DOFs a = warp, warp+16, ... of the row chunk.  Stage 1, lanes = column cells c2 (two per lane):
// v(c2) = sum_{(c1,k1) of a} +-q1[c1][k1] T[c1][c2]; the three products q2[c2][k].v(c2) (cell c2's contribution to
// its vertices) go to a per-warp scratch.  Stage 2, lanes = column DOFs b (ascending reference ids: neighbouring lanes
// touch neighbouring addresses of row a): sum of the scratch entries of b's cells, added to the old value (loaded
// before stage 1) and stored.  Transposed entries (tiles whose column patch's rows are owned and not left to the
// symmetrisation pass) get the same value with one store per lane.  Every entry is owned by this CTA: plain loads
// and stores, no atomics, no block-wide barriers.
// role_sel 0: all entries; 1: only entries with a <= b (a second-role pass follows); 2: only a > b
__device__ __forceinline__ void drain_pass(Smem& S, const LmatArgs& A, const ChunkState& I, const ChunkState& J, int flags, int role_sel,  // @region D_drain
                                           int tid) {
  const double* __restrict__ T = S.T;
  const bool diag = flags & 1, mirror = A.self && (flags & 3) && !(flags & 8);
  const int ndI = I.ndof, ndJ = J.ndof;
  const int lane = tid & 31, warp = tid >> 5;
  double* __restrict__ P = S.u.P[warp];  // [3][kCH] products of this warp
  const int lim = A.fast_lim;            // 64 (tests of the generic path: 32 or 0)
  // this lane's column DOFs b = lane, lane+32: reference id, output column, incidence list (scratch index k*64+cell = low 8
  // bits of the incidence code, bit 8 = negative) packed in registers (longer lists: generic path)
  constexpr int MI = 8;
  int ninc[2], ob[2], oc[2], rb[2];
  unsigned codes[2][MI / 2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int ib = lane + 32 * r;
    ninc[r] = -1;  // -1: not handled here
    ob[r] = 0;
    oc[r] = -1;
    rb[r] = -1;
#pragma unroll
    for (int i = 0; i < MI / 2; i++) codes[r][i] = 0;
    if (ib < ndJ && ib < lim) {
      const int i0 = J.x.iptr[ib];
      const int n = J.x.iptr[ib + 1] - i0;
      if (n <= MI) {
        ninc[r] = n;
#pragma unroll
        for (int i = 0; i < MI; i++)
          if (i < n) codes[r][i >> 1] |= (unsigned)J.x.inc[i0 + i] << (16 * (i & 1));
      } else {
        ninc[r] = MI + 1;  // long list: summed from shared memory
      }
      ob[r] = J.x.orig[ib];
      oc[r] = A.col_map ? A.col_map[ob[r]] : ob[r];
      rb[r] = mirror ? J.row[ib] : -1;
    }
  }
  auto wanted = [&](int oa, int obv) {  // role rule of entry (a, b)
    if (!A.self) return true;
    const bool role1 = oa <= obv;
    return !((diag && !role1) || (role_sel == 1 && !role1) || (role_sel == 2 && role1));
  };
  auto col_sum = [&](int ib) {  // generic: contribution to column DOF ib from the scratch
    double acc = 0.0;
    for (int i = J.x.iptr[ib]; i < J.x.iptr[ib + 1]; i++) {
      const unsigned w = J.x.inc[i];
      const double v = P[w & 255u];
      acc += (w & 256u) ? -v : v;
    }
    return acc;
  };
  // the basis vectors of the lane's two column cells stay in registers across the row DOFs of the pass (18 of the ~70
  // shared-memory loads per row DOF; 100k-vessel shard: 359.6 -> 356.6 ms)
  double q2a[9], q2b[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    q2a[k] = J.g[(10 + k) * kCH + lane];
    q2b[k] = J.g[(10 + k) * kCH + lane + 32];
  }
#pragma unroll 1
  for (int ia = warp; ia < ndI; ia += NW) {
    const int ra = I.row[ia];
    if (ra < 0 && !mirror) continue;
    const int oa = I.x.orig[ia];
    double* const rowp = A.out + (long long)(ra >= 0 ? ra : 0) * A.ld;
    // old values of the direct entries
    double old[2];
    bool use[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      use[r] = ninc[r] >= 0 && wanted(oa, ob[r]);
      old[r] = 0.0;
      if (use[r] && ra >= 0 && oc[r] >= 0) old[r] = __ldcg(rowp + oc[r]);
    }
    // stage 1
    double vx0 = 0.0, vy0 = 0.0, vz0 = 0.0, vx1 = 0.0, vy1 = 0.0, vz1 = 0.0;
    const int i1e = I.x.iptr[ia + 1];
#pragma unroll 4
    for (int i1 = I.x.iptr[ia]; i1 < i1e; i1++) {
      const unsigned w1 = I.x.inc[i1];
      const int c1 = w1 & 63, k1 = (w1 >> 6) & 3;
      double t0 = T[c1 * TS + lane], t1 = T[c1 * TS + lane + 32];
      if (w1 & 256) {
        t0 = -t0;
        t1 = -t1;
      }
      const double qx = I.g[(10 + 3 * k1) * kCH + c1], qy = I.g[(11 + 3 * k1) * kCH + c1], qz = I.g[(12 + 3 * k1) * kCH + c1];
      vx0 = fma(qx, t0, vx0);
      vy0 = fma(qy, t0, vy0);
      vz0 = fma(qz, t0, vz0);
      vx1 = fma(qx, t1, vx1);
      vy1 = fma(qy, t1, vy1);
      vz1 = fma(qz, t1, vz1);
    }
    __syncwarp();  // stage 2 of the previous row DOF is done with the scratch
#pragma unroll
    for (int k = 0; k < 3; k++) {
      P[k * kCH + lane] = fma(q2a[3 * k + 2], vz0, fma(q2a[3 * k + 1], vy0, q2a[3 * k] * vx0));
      P[k * kCH + lane + 32] = fma(q2b[3 * k + 2], vz1, fma(q2b[3 * k + 1], vy1, q2b[3 * k] * vx1));
    }
    __syncwarp();
    // stage 2
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (!use[r]) continue;
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
        acc = col_sum(lane + 32 * r);
      }
      if (ra >= 0 && oc[r] >= 0) __stcg(rowp + oc[r], fma(acc, A.scale, old[r]));  // (every path: one fused scale-and-add)
      if (rb[r] >= 0 && oa != ob[r]) {
        double* pm = A.out + (long long)rb[r] * A.ld + oa;
        __stcg(pm, fma(acc, A.scale, __ldcg(pm)));
      }
    }
    // column DOFs beyond the register-held ones (more than 64 per chunk: rare)
#pragma unroll 1
    for (int ib = lane + min(lim, 64); ib < ndJ; ib += 32) {
      const int obv = J.x.orig[ib];
      if (!wanted(oa, obv)) continue;
      const double acc = col_sum(ib);
      const int ocv = A.col_map ? A.col_map[obv] : obv;
      if (ra >= 0 && ocv >= 0) __stcg(rowp + ocv, fma(acc, A.scale, __ldcg(rowp + ocv)));
      if (mirror && oa != obv) {
        const int rbv = J.row[ib];
        if (rbv >= 0) {
          double* pm = A.out + (long long)rbv * A.ld + oa;
          __stcg(pm, fma(acc, A.scale, __ldcg(pm)));
        }
      }
    }
  }
}

// ---- streamed build: mirror tasks and band bookkeeping ---------------------------------------------------------------
// The matrix lies in the reference layout (out[row DOF][column DOF], both reference ids); a band owns the rows [R0,R1) and
// its tiles filled, of every pair {i,k} with i in the band and k >= R0, the entry [i][k] iff patch(i) < patch(k), or the
// patches are equal and i <= k -- else [k][i] (then k lies in the band too).  The other entry of the pair is its copy
// (thin_wall.F90:1146-1151): afterwards the rows of the band are final from column R0 on, and so are the columns [R0,R1)
// of all later rows -- the L-shaped part of the matrix that leaves for the host while the later bands are evaluated.
// One 32 x 32 block of that pass, handled by one warp through its own 32 x 33 scratch (coalesced both ways): rows
// i = ib.., columns k = kb..
__device__ __forceinline__ void mirror_block(double* __restrict__ buf, int R1, int N, int ib, int kb, bool diag, const int* __restrict__ ref_patch,  // @region mirror_task
                                             double* __restrict__ out, long long ld, int lane) {
  const int pi = ib + lane < R1 ? ref_patch[ib + lane] : -1, pk = kb + lane < N ? ref_patch[kb + lane] : -1;
#pragma unroll 8
  for (int ii = 0; ii < 32; ii++)  // A[ii][kk] = out[i][k], lanes over k
    if (ib + ii < R1 && kb + lane < N) buf[ii * 33 + lane] = __ldcg(out + (long long)(ib + ii) * ld + kb + lane);
  __syncwarp();
#pragma unroll 8
  for (int kk = 0; kk < 32; kk++) {  // out[k][i] = out[i][k] where [i][k] was evaluated, lanes over i
    const int k = kb + kk, i = ib + lane, pkk = __shfl_sync(0xffffffffu, pk, kk);
    if (k < N && i < R1 && i != k && (pi < pkk || (pi == pkk && i < k))) out[(long long)k * ld + i] = buf[lane * 33 + kk];
  }
  if (!diag && kb < R1) {  // rows k of the band: entries they hold may belong to the rows i
    __syncwarp();
#pragma unroll 8
    for (int kk = 0; kk < 32; kk++)  // Bt[kk][ii] = out[k][i], lanes over i
      if (kb + kk < R1 && ib + lane < R1) buf[kk * 33 + lane] = __ldcg(out + (long long)(kb + kk) * ld + ib + lane);
    __syncwarp();
#pragma unroll 8
    for (int ii = 0; ii < 32; ii++) {  // out[i][k] = out[k][i] where [k][i] was evaluated, lanes over k
      const int i = ib + ii, k = kb + lane, pii = __shfl_sync(0xffffffffu, pi, ii);
      if (i < R1 && k < R1 && (pk < pii || (pk == pii && k < i))) out[(long long)i * ld + k] = buf[lane * 33 + ii];
    }
  }
  __syncwarp();
}
__device__ __forceinline__ int mirror_task_count(int R0, int R1, int N) {
  return ((R1 - R0 + 31) / 32) * (((N - R0 + 31) / 32 + kMirKC - 1) / kMirKC);
}
// mirror task `task` of the band with rows [R0,R1): 32 rows of the band against kMirKC column blocks, one block per warp at
// a time (scalar arguments: a reference to the kernel's parameter block would force a local copy of it)
__device__ __noinline__ void mirror_task(double* __restrict__ scratch, int R0, int R1, int N, const int* __restrict__ ref_patch, double* __restrict__ out,
                                         long long ld, int task, int tid) {
  const int nbk = (N - R0 + 31) / 32, nkc = (nbk + kMirKC - 1) / kMirKC;
  const int bx = task / nkc, kc = task - bx * nkc;
  double* buf = scratch + (tid >> 5) * (32 * 33);
  const int by1 = min(nbk, (kc + 1) * kMirKC);
  for (int by = max(bx, kc * kMirKC) + (tid >> 5); by < by1; by += NW)
    mirror_block(buf, R1, N, R0 + 32 * bx, R0 + 32 * by, by == bx, ref_patch, out, ld, tid & 31);
}
// next work item of a streamed build (one thread): a mirror task of a band whose tiles are all done, else the next tile;
// when the tiles are used up, the remaining bands are waited for (their tiles are running on other CTAs).
// item[0]: tile, -1: exit, -2: mirror task item[1] of band item[2]; item[3]: cursor over the bands, item[4]: band of the
// tile this CTA just finished (or -1)
__device__ __noinline__ void fetch_banded(int* __restrict__ item, int* tile_counter, int ntiles, int nbands, const int* __restrict__ tile_band,
                                          const int* __restrict__ band_ntiles, const int* __restrict__ band_ref, int N, int* band_state) {
  int* tiles_done = band_state;
  int* mir_next = band_state + nbands;
  if (item[4] >= 0) atomicAdd(&tiles_done[item[4]], 1);  // (all threads fenced their stores before the barrier)
  item[4] = -1;
  bool tiles_left = true;
  for (;;) {
    for (int b = item[3]; b < nbands; b++) {
      if (*(volatile int*)&tiles_done[b] < band_ntiles[b]) break;  // bands hand out their mirror tasks in order
      const int k = atomicAdd(&mir_next[b], 1);
      if (k < mirror_task_count(band_ref[2 * b], band_ref[2 * b + 1], N)) {
        __threadfence();  // the tile stores of the other CTAs (released by their fences) before this task's loads
        item[0] = -2;
        item[1] = k;
        item[2] = b;
        return;
      }
      item[3] = b + 1;
    }
    if (item[3] >= nbands) {
      item[0] = -1;
      return;
    }
    if (tiles_left) {
      const int t = atomicAdd(tile_counter, 1);
      if (t < ntiles) {
        item[4] = tile_band[t];
        item[0] = t;
        return;
      }
      tiles_left = false;
    }
    __nanosleep(2000);
  }
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
  __syncthreads();
  uint32_t phI = 0, phJ = 0;  // mbarrier phases: row slot, bit s = column slot s
  unsigned long long st_far = 0, st_near = 0, st_eval = 0, st_phi = 0;
  if (A.stats && tid == 0 && blockIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    A.stats[6] = t0;
  }

  if (tid == 0) {
    S.item[3] = 0;
    S.item[4] = -1;
  }
  for (;;) {
    if (A.nbands) __threadfence();  // streamed build: this tile's stores before its band's count
    __syncthreads();  // the previous tile is drained (chunk slots, T and the scratch are free)
    if (tid == 0) {
      if (A.nbands) fetch_banded(S.item, A.tile_counter, A.ntiles, A.nbands, A.tile_band, A.band_ntiles, A.band_ref, A.N, A.band_state);
      else S.item[0] = atomicAdd(A.tile_counter, 1);
    }
    __syncthreads();
    const int t = S.item[0];
    if (t >= A.ntiles || t == -1) break;
    if (t == -2) {  // mirror task of a finished band
      const int b = S.item[2], R0 = A.band_ref[2 * b], R1 = A.band_ref[2 * b + 1];
      mirror_task(S.T, R0, R1, A.N, A.ref_patch, A.out, A.ld, S.item[1], tid);
      __threadfence_system();
      __syncthreads();
      if (tid == 0) {
        int* mir_done = A.band_state + 2 * A.nbands;
        if (atomicAdd(&mir_done[b], 1) == mirror_task_count(R0, R1, A.N) - 1) {  // the band's part of the matrix is final
          __threadfence_system();
          A.host_flags[b] = 1;
          __threadfence_system();
        }
      }
      continue;
    }
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
    for (int ci = ci0; ci < ci1; ci++) {  // @region chunk_loop
      mbar_wait(&S.I.bar, phI);
      phI ^= 1;
      const ChunkState& I = S.I;
      const double ox = I.cx, oy = I.cy, oz = I.cz;
      // row side of the sweep (first used after the barrier of the first pass)
      prep_chunk(I, S.vfI, S.flI, S.cenI, S.tabI, true, ox, oy, oz, tid);
      bool tabI_dirty = false;
      for (int cj = cj0; cj < cj1; cj++) {
        const ChunkState& J = S.J[js];
        mbar_wait(&S.J[js].bar, (phJ >> js) & 1u);
        phJ ^= 1u << js;
        // (everybody is done with the previous pass: barrier at its end)
        // prefetch the next column chunk (its slot was used by the previous pass)
        const bool lastj = cj + 1 == cj1, lasti = ci + 1 == ci1;
        if (tid == 0 && !(lastj && lasti)) stage_issue(S.J[js ^ 1], 1, A, lastj ? cj0 : cj + 1);
        prep_chunk(J, S.vfJ, S.flJ, S.cenJ, S.u.tabJ, false, ox, oy, oz, tid);
        if (tabI_dirty) {  // the previous pass used the row tables' memory for a larger rule
          build_table(S.tabI, CI, I.g, 0, I.ncell, 4, 6, ox, oy, oz, true, tid, NT);
          build_table(S.tabI + 6 * 2 * CI, CI, I.g, 0, I.ncell, 5, 7, ox, oy, oz, true, tid, NT);
          build_table(S.tabI + 13 * 2 * CI, CI, I.g, 0, I.ncell, 6, 12, ox, oy, oz, true, tid, NT);
          tabI_dirty = false;
        }
        if (tid < NCLS) {
          S.pb.cnt[tid] = 0;
          S.pb.pos[tid] = 0;
        }
        if (tid < kTabClsMax + 1) S.pb.nleft[tid] = 0;
        if (tid == 0) {
          S.pb.rowhead = 0;
          S.pb.lhead = 0;
          S.pb.qhead = 0;
          S.pb.both_count = 0;
        }
        __syncthreads();  // tables, FP32 copies and counters of the pass
        sweep_rows(S, A, I, J, flags, tid, st_far, st_eval);
        __syncthreads();  // T of the full batches, partial batches, pair codes and class counts complete
        int nother = 0;
#pragma unroll
        for (int c = kTabClsMax + 1; c < NCLS; c++) nother += S.pb.cnt[c];
        if (nother) bin_others(S, I, tid);
        eval_leftovers(S, A, I, J, tid);
        if (nother) {
          __syncthreads();  // bins and queue of the other classes; everybody is done with the tables of the sweep
          const int ngrp = S.pb.ngrp;
#pragma unroll 1
          for (int g = 0; g < ngrp; g++) {
            const int tc = S.pb.gcls[g];
            if (g > 0) {
              __syncthreads();  // previous group done with its tables and the queue head
              if (tid == 0) S.pb.qhead = 0;
            }
            if (tc >= 0) {  // point tables of this rule: row side over the sweep's row tables (rebuilt before the next pass)
              build_table(S.tabI, CI, I.g, 0, I.ncell, tc + 4, c_cls_np[tc], ox, oy, oz, true, tid, NT);
              build_table(S.u.tabJ, kCH, J.g, 0, J.ncell, tc + 4, c_cls_np[tc], ox, oy, oz, false, tid, NT);
              tabI_dirty = true;
              __syncthreads();
            }
            eval_others(S, A, I, J, tid, g, st_near, st_phi);
          }
        }
        const bool two_pass = S.pb.both_count > 0;
        __syncthreads();  // T complete; the column tables are dead: their memory is the contraction scratch
        if (!(A.debug_skip & 4)) drain_pass(S, A, I, J, flags, two_pass ? 1 : 0, tid);
        if (two_pass) {
          __syncthreads();  // first contraction done with T
          if (tid == 0) S.pb.qhead = 0;
          __syncthreads();
          eval_role2(S, A, I, J, tid, st_near, st_phi);
          __syncthreads();
          if (!(A.debug_skip & 4)) drain_pass(S, A, I, J, flags, 2, tid);
        }
        __syncthreads();  // pass done: T, lists, scratch and the other column slot are free
        js ^= 1;
      }
      if (ci + 1 < ci1) {
        if (tid == 0) stage_issue(S.I, 0, A, ci + 1);  // (the barrier at the end of the last pass: nobody reads the row chunk)
        __syncthreads();  // chunk header
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
// i in [i0,i1) (rows of dst, row index i-i0) and j in [j0,j1) (rows of src, row index j-j0):
// dst[i][orig(j)] = src[j][orig(i)]  (thin_wall.F90:1146-1151 across shards).  checker = 0: every entry (src holds the
// whole block: bands of one device); 1: only the entries of the tiles the OTHER shard evaluated (symmetric shards: tile
// {patch(i), patch(j)} belongs to the shard of j iff !sym_tile_is_mine(patch(i), patch(j))) -- the rest of dst's block was
// computed in place, and the corresponding part of src is what src's owner is copying from here at the same time.
constexpr int kCrossI = 4;  // 32-row sub-tiles of dst a block handles against the same 32 rows of src (page reuse on the peer)
__global__ void symmetrize_cross_kernel(int i0, int i1, int j0, int j1, const int* __restrict__ dof_orig, const int* __restrict__ dof_patch,
                                        int checker, double* __restrict__ dst, const double* __restrict__ src, long long ld) {
  __shared__ double tile[32][33];
  __shared__ int oi_s[32], oj_s[32], pi_s[32], pj_s[32];
  const int jb = j0 + 32 * blockIdx.y;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (ty == 1) {
    oj_s[tx] = jb + tx < j1 ? dof_orig[jb + tx] : -1;
    pj_s[tx] = jb + tx < j1 ? dof_patch[jb + tx] : -1;
  }
#pragma unroll 1
  for (int sub = 0; sub < kCrossI; sub++) {
    const int ib = i0 + 32 * (blockIdx.x * kCrossI + sub);
    if (ib >= i1) break;
    __syncthreads();  // previous sub-tile done with the shared arrays
    if (ty == 0) {
      oi_s[tx] = ib + tx < i1 ? dof_orig[ib + tx] : -1;
      pi_s[tx] = ib + tx < i1 ? dof_patch[ib + tx] : -1;
    }
    __syncthreads();
    if (checker) {  // block-uniform skip: the sub-tile lies in one patch pair whose tile is mine (patches are contiguous ranges)
      const int il = min(31, i1 - 1 - ib), jl = min(31, j1 - 1 - jb);
      if (pi_s[0] == pi_s[il] && pj_s[0] == pj_s[jl] && tw::sym_tile_is_mine(pi_s[0], pj_s[0])) continue;
    }
    for (int jj = ty; jj < 32; jj += 8) {  // read src[j][orig(i)], lanes over i
      const int j = jb + jj;
      if (j < j1 && oi_s[tx] >= 0 && !(checker && tw::sym_tile_is_mine(pi_s[tx], pj_s[jj])))
        tile[jj][tx] = __ldcg(src + (long long)(j - j0) * ld + oi_s[tx]);
    }
    __syncthreads();
    for (int ii = ty; ii < 32; ii += 8) {  // write dst[i][orig(j)], lanes over j
      const int i = ib + ii;
      if (i < i1 && oj_s[tx] >= 0 && !(checker && tw::sym_tile_is_mine(pi_s[ii], pj_s[tx]))) dst[(long long)(i - i0) * ld + oj_s[tx]] = tile[tx][ii];
    }
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
static std::string upload(const std::vector<T>& h, T** d, size_t* cap = nullptr) {
  const size_t n = std::max<size_t>(h.size(), 1) * sizeof(T);
  if (!(cap && *d && *cap >= n)) {
    if (cap && *d) cudaFree(*d);
    *d = nullptr;
    if (cap) *cap = 0;
    CK(cudaMalloc((void**)d, n));
    if (cap) *cap = n;
  }
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
  std::string e;
  if (!(e = upload(ps.chunks, &chunks, &cap[0])).empty()) return e;
  if (!(e = upload(ps.geom, &geom, &cap[1])).empty()) return e;
  if (!(e = upload(ps.cell_dmin, &dmin, &cap[2])).empty()) return e;
  if (!(e = upload(ps.cell_dmax, &dmax, &cap[3])).empty()) return e;
  if (!(e = upload(ps.chunk_dof, &chunk_dof, &cap[4])).empty()) return e;
  if (!(e = upload(ps.chunk_inc_ptr, &inc_ptr, &cap[5])).empty()) return e;
  if (!(e = upload(ps.inc, &inc, &cap[6])).empty()) return e;
  if (!(e = upload(ps.patch_chunk_ptr, &patch_chunk_ptr, &cap[7])).empty()) return e;
  if (!(e = upload(ps.dof_orig, &dof_orig, &cap[8])).empty()) return e;
  {
    std::vector<int> dp(ps.ndof, 0);
    for (int p = 0; p < ps.npatch; p++)
      for (int i = ps.patch_dof_ptr[p]; i < ps.patch_dof_ptr[p + 1]; i++) dp[i] = p;
    if (!(e = upload(dp, &dof_patch, &cap[9])).empty()) return e;
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
    if (!(e = upload(ax, &aux, &cap[10])).empty()) return e;
    nchunk = ps.nchunk;
  }
  bytes = ps.chunks.size() * sizeof(ChunkMeta) + ps.geom.size() * 8 + (ps.cell_dmin.size() + ps.cell_dmax.size()) * 4 +
          (ps.chunk_dof.size() + ps.chunk_inc_ptr.size() + ps.patch_chunk_ptr.size() + ps.dof_orig.size()) * 4 + ps.inc.size() * 2;
  return "";
}
void DevicePatchSet::release() {
  for (size_t& c : cap) c = 0;
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
                                 cudaStream_t stream, bool checker) {
  if (i1 <= i0 || j1 <= j0) return "";
  twk::symmetrize_cross_kernel<<<dim3((i1 - i0 + 32 * twk::kCrossI - 1) / (32 * twk::kCrossI), (j1 - j0 + 31) / 32), dim3(32, 8), 0, stream>>>(i0, i1, j0, j1, A.dof_orig, A.dof_patch,
                                                                                                      checker ? 1 : 0, dst, src, ld);
  CK(cudaGetLastError());
  note_launch();
  return "";
}

std::string gpu_lmat_tiles(const DevicePatchSet& A, const DevicePatchSet& B, const std::vector<Tile>& tiles,
                           const std::vector<int>& row_out, bool self, double* d_out, long long ld, cudaStream_t stream,
                           unsigned long long* h_stats, const int* d_col_map, bool symmetrize, const StreamBands* sb) {
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
  a.nbands = 0;
  a.tile_band = a.band_ntiles = a.band_ref = a.ref_patch = nullptr;
  a.band_state = nullptr;
  a.host_flags = nullptr;
  a.N = 0;
  if (sb && sb->nbands > 0) {
    const int nb = sb->nbands;
    int* d_flags = nullptr;
    CK(cudaHostGetDevicePointer((void**)&d_flags, sb->flags, 0));
    a.nbands = nb;
    a.band_ntiles = sb->d_bands;
    a.band_ref = sb->d_bands + nb;
    a.band_state = sb->d_bands + 3 * nb;
    a.tile_band = sb->d_tile_band;
    a.ref_patch = sb->d_ref_patch;
    a.N = sb->N;
    a.host_flags = d_flags;
  }
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
