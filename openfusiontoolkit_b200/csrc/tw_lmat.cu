// tw_lmat.cu -- dense element<->element inductance build on sm_100a (FP64, no tensor cores).
//
// Replaces the O(nc^2) OpenMP loop nest of tw_compute_LmatDirect (src/physics/thin_wall.F90:
// 1008-1126) with an owner-computes tiling: one CTA owns the output tile (row patch x column
// patch), stages a chunk of "row" triangles and a chunk of "column" triangles in shared memory
// (1-D bulk async copies of the contiguous SoA chunk records, mbarrier-tracked), evaluates the
// pair integrals T(c1,c2) for the 64x64 chunk pair into a shared-memory tile, and contracts
// them onto the vertex/hole DOFs.  Every L entry is written by exactly one thread of exactly
// one CTA (plain read-modify-write between barriers): no atomics, deterministic summation.
//
// Role rule (SURVEY hard part 1): for entry (a,b) with a<=b in reference numbering the cell
// carrying `a` is the analytic side of near pairs.  T1 = T(c1 analytic), T2 = T(c2 analytic).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "quad_tables.h"
#include "tw_device.cuh"
#include "tw_gpu.h"

namespace twk {

using tw::kCH;
using tw::kGeomRows;
constexpr int NT = 512;            // threads per CTA (16 warps), one CTA per SM
constexpr int NW = NT / 32;
constexpr int kMaxNear = kCH * kCH;

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
  unsigned long long* stats;          // [0] far pairs, [1] near T evaluations, [2] 1/r evaluations, [3] phipot evals
};

struct Smem {
  double gI[kGeomRows * kCH];
  double gJ[kGeomRows * kCH];
  double nI[3 * kCH];   // unit normals of row cells
  double nJ[3 * kCH];
  double T1[kCH * kCH];  // [c1][c2]
  double T2[kCH * kCH];
  unsigned int near_list[kMaxNear];
  int dminI[kCH], dmaxI[kCH], dminJ[kCH], dmaxJ[kCH];
  int dofI[tw::kMaxChunkDof], dofJ[tw::kMaxChunkDof];
  int iptrI[tw::kMaxChunkDof + 1], iptrJ[tw::kMaxChunkDof + 1];
  uint16_t incI[tw::kMaxChunkInc], incJ[tw::kMaxChunkInc];
  unsigned long long bar[2];
  int near_count;
  int tile_id;
};

// ---- far-field tensor quadrature -----------------------------------------------------------
// T = area_i area_j sum_p sum_q w_p w_q / |x_p(i) - x_q(j)|, same rule on both triangles
// (thin_wall.F90:1069-1083).  j-side points are held in registers in blocks of <= 8; the i-side
// point is recomputed per p (warp-uniform broadcast from shared memory).
template <int N>
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

__device__ __forceinline__ double far_dispatch(const double* gI, int c1, const double* gJ, int c2, int iquad) {
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

// near pair: T = area_q * sum_q w_q phi_{tri A}(x_q(tri Q)); lanes parallelise over q
// (thin_wall.F90:1061-1068).  gA/cA = analytic triangle, gQ/cQ = quadrature triangle.
__device__ __forceinline__ double near_pair(const double* gA, const double* nA, int cA, const double* gQ, int cQ,
                                            int iquad, int lane) {
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
  for (int q = lane; q < n; q += 32) {
    double b0 = bp[3 * q], b1 = bp[3 * q + 1], b2 = bp[3 * q + 2];
    double x = xquad(b0, b1, b2, PQ[0], PQ[3], PQ[6]);
    double y = xquad(b0, b1, b2, PQ[1], PQ[4], PQ[7]);
    double z = xquad(b0, b1, b2, PQ[2], PQ[5], PQ[8]);
    s += bw[q] * phipot(PA, nh, x, y, z);
  }
  s = warp_sum(s);
  return s * gQ[9 * kCH + cQ];
}

__device__ __forceinline__ int classify_pair(const double* gI, int c1, const double* gJ, int c2) {
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

__device__ __forceinline__ void load_chunk(Smem& S, int side, const LmatArgs& A, int chunk, unsigned long long* bar,
                                           uint32_t& phase) {
  // side 0: row chunk (I), 1: column chunk (J).  Geometry record via one bulk async copy issued
  // by a single thread; the small index lists by all threads; normals computed after arrival.
  const tw::ChunkMeta* cms = side ? A.chunksB : A.chunksA;
  const double* geom = side ? A.geomB : A.geomA;
  double* g = side ? S.gJ : S.gI;
  const tw::ChunkMeta cm = cms[chunk];
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order prior generic accesses before the async write
    mbar_expect_tx(bar, kGeomRows * kCH * 8);
    bulk_g2s(g, geom + (size_t)chunk * kGeomRows * kCH, kGeomRows * kCH * 8, bar);
  }
  const int* dmn = (side ? A.dminB : A.dminA) + (size_t)chunk * kCH;
  const int* dmx = (side ? A.dmaxB : A.dmaxA) + (size_t)chunk * kCH;
  const int* cdof = (side ? A.chunk_dofB : A.chunk_dofA) + cm.dof_off;
  const int* iptr = (side ? A.inc_ptrB : A.inc_ptrA) + cm.dof_off + chunk;
  const uint16_t* inc = (side ? A.incB : A.incA) + cm.inc_off;
  int* sdmn = side ? S.dminJ : S.dminI;
  int* sdmx = side ? S.dmaxJ : S.dmaxI;
  int* sdof = side ? S.dofJ : S.dofI;
  int* sptr = side ? S.iptrJ : S.iptrI;
  uint16_t* sinc = side ? S.incJ : S.incI;
  for (int i = threadIdx.x; i < kCH; i += NT) {
    sdmn[i] = dmn[i];
    sdmx[i] = dmx[i];
  }
  for (int i = threadIdx.x; i < cm.ndof; i += NT) sdof[i] = cdof[i];
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

__global__ void __launch_bounds__(NT, 1) lmat_tile_kernel(const LmatArgs A) {
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

    for (int ci = ci0; ci < ci1; ci++) {
      __syncthreads();  // previous contraction finished with gI lists
      load_chunk(S, 0, A, ci, &S.bar[0], phI);
      const int ncI = A.chunksA[ci].ncell, ndI = A.chunksA[ci].ndof;
      for (int cj = cj0; cj < cj1; cj++) {
        __syncthreads();  // previous contraction finished with gJ / T tiles
        load_chunk(S, 1, A, cj, &S.bar[1], phJ);
        const int ncJ = A.chunksB[cj].ncell, ndJ = A.chunksB[cj].ndof;
        if (tid == 0) S.near_count = 0;
        __syncthreads();

        // ---------------- phase 1: classification + far field --------------------------------
        // warp w: column half (w&1), rows (w>>1) + 8m
        {
          const int c2 = lane + 32 * (warp & 1);
          for (int m = 0; m < kCH / 8; m++) {
            const int c1 = (warp >> 1) + 8 * m;
            if (c1 >= ncI) break;  // warp-uniform
            int iq = 0;
            bool n1 = false, n2 = false;
            if (c2 < ncJ) {
              if (A.self) {
                n1 = S.dminI[c1] <= S.dmaxJ[c2];
                n2 = want2 && (S.dmaxI[c1] > S.dminJ[c2]);
                if (diag) n2 = false;
              } else {
                n1 = true;
              }
              if (n1 || n2) iq = classify_pair(S.gI, c1, S.gJ, c2);
            }
            double tval = 0.0;
            if (iq > 10) {
              int k = atomicAdd(&S.near_count, 1);
              S.near_list[k] = (unsigned)c1 | ((unsigned)c2 << 6) | ((unsigned)iq << 12) | ((unsigned)n1 << 17) |
                               ((unsigned)n2 << 18);
            }
            unsigned todo = __ballot_sync(0xffffffffu, iq >= 4 && iq <= 10);
            while (todo) {
              const int leader = __ffs(todo) - 1;
              const int r = __shfl_sync(0xffffffffu, iq, leader);
              const bool mine = (iq == r);
              if (mine) {
                tval = far_dispatch(S.gI, c1, S.gJ, c2, r);
                st_far++;
                st_eval += (unsigned long long)c_qnp[r] * c_qnp[r];
              }
              todo &= ~__ballot_sync(0xffffffffu, mine);
            }
            if (c2 < kCH) {
              S.T1[c1 * kCH + c2] = tval;
              S.T2[c1 * kCH + c2] = tval;
            }
          }
        }
        __syncthreads();
        // ---------------- phase 2: near field (one warp per pair, lanes over points) ----------
        {
          const int nn = S.near_count;
          for (int k = warp; k < nn; k += NW) {
            const unsigned e = S.near_list[k];
            const int c1 = e & 63, c2 = (e >> 6) & 63, iq = (e >> 12) & 31;
            if (e & (1u << 17)) {
              double v = near_pair(S.gI, S.nI, c1, S.gJ, c2, iq, lane);
              if (lane == 0) S.T1[c1 * kCH + c2] = v;
              st_near += (lane == 0);
              st_phi += (lane == 0) ? c_qnp[iq] : 0;
            }
            if (e & (1u << 18)) {
              double v = near_pair(S.gJ, S.nJ, c2, S.gI, c1, iq, lane);
              if (lane == 0) S.T2[c1 * kCH + c2] = v;
              st_near += (lane == 0);
              st_phi += (lane == 0) ? c_qnp[iq] : 0;
            }
          }
        }
        __syncthreads();
        // ---------------- phase 3: contraction onto DOFs, owner writes -------------------------
        {
          const int nent = ndI * ndJ;
          for (int e = tid; e < nent; e += NT) {
            const int ia = e / ndJ, ib = e - ia * ndJ;
            const int da = S.dofI[ia], db = S.dofJ[ib];
            const int oa = A.dof_origA[da], ob = A.dof_origB[db];
            bool role1 = true;
            if (A.self) {
              role1 = (oa <= ob);
              if (diag && !role1) continue;
            }
            const double* T = role1 ? S.T1 : S.T2;
            double acc = 0.0;
            for (int i1 = S.iptrI[ia]; i1 < S.iptrI[ia + 1]; i1++) {
              const unsigned w1 = S.incI[i1];
              const int c1 = w1 & 63, k1 = (w1 >> 6) & 3;
              const double e1x = S.gI[(10 + 3 * k1) * kCH + c1], e1y = S.gI[(11 + 3 * k1) * kCH + c1],
                           e1z = S.gI[(12 + 3 * k1) * kCH + c1];
              double ux = 0.0, uy = 0.0, uz = 0.0;
              for (int i2 = S.iptrJ[ib]; i2 < S.iptrJ[ib + 1]; i2++) {
                const unsigned w2 = S.incJ[i2];
                const int c2 = w2 & 63, k2 = (w2 >> 6) & 3;
                double tv = T[c1 * kCH + c2];
                if (w2 & 256) tv = -tv;
                ux = fma(S.gJ[(10 + 3 * k2) * kCH + c2], tv, ux);
                uy = fma(S.gJ[(11 + 3 * k2) * kCH + c2], tv, uy);
                uz = fma(S.gJ[(12 + 3 * k2) * kCH + c2], tv, uz);
              }
              double dsum = e1x * ux + e1y * uy + e1z * uz;
              acc += (w1 & 256) ? -dsum : dsum;
            }
            acc *= A.scale;
            const int ra = A.row_out[da];
            if (ra >= 0) A.out[(long long)ra * A.ld + ob] += acc;
            if (A.self && (mirror || diag) && oa != ob) {
              const int rb = A.row_out[db];
              if (rb >= 0) A.out[(long long)rb * A.ld + oa] += acc;
            }
          }
        }
      }
    }
  }
  if (A.stats) {
    st_far = warp_sum((double)st_far);  // counts are < 2^53
    st_near = warp_sum((double)st_near);
    st_eval = warp_sum((double)st_eval);
    st_phi = warp_sum((double)st_phi);
    if (lane == 0) {
      atomicAdd(&A.stats[0], st_far);
      atomicAdd(&A.stats[1], st_near);
      atomicAdd(&A.stats[2], st_eval);
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
  for (int k = 5; k <= 18; k++) {
    thr[k - 5] = order_threshold(k);
    thr2[k - 5] = thr[k - 5] * thr[k - 5];
  }
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
  a.stats = d_stats;
  int dev = 0, nsm = 148;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  int grid = (int)std::min<size_t>(tiles.size(), (size_t)nsm);
  twk::lmat_tile_kernel<<<grid, twk::NT, sizeof(twk::Smem), stream>>>(a);
  CK(cudaGetLastError());
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
