// tw_ops.cu -- coil / sensor coupling, B-field reconstruction operator, mutual inductance,
// statistics and operator caches.  (Included by tw_unity.cu after tw_lmat.cu.)
//
//  elem_filament_kernel   : cell x polyline potentials + trapezoid   thin_wall.F90:649-707, 1512-1561
//  dof_gather_kernel      : owner-computes scatter to vertex/hole DOFs              :708-723, 1562-1577
//  filament_mutual_kernel : filament<->filament Neumann sums                       :814-849, 1597-1632
//  bel_tile_kernel        : element -> B(vertex) operator                         :2017-2110
//  filament_bfield_kernel : midpoint Biot-Savart of filaments at the vertices     :2119-2168
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tw_device.cuh"
#include "tw_gpu.h"
#include "tw_ops.h"

namespace twk {

// ---------------------------------------------------------------------------------------------
// potential of one cell at one point: analytic if iquad>10 else one-sided quadrature
// (thin_wall.F90:666-694 / :1528-1552)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double cell_point_pot(const double* P, double area, const double* nhat, double x, double y,
                                                 double z, double* d2min_out) {
  double d2min = 1.e300, d2max = 0.0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    double dx = P[3 * a] - x, dy = P[3 * a + 1] - y, dz = P[3 * a + 2] - z;
    double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
    d2min = fmin(d2min, d2);
    d2max = fmax(d2max, d2);
  }
  if (d2min_out) *d2min_out = d2min;
  double floor2 = area * 2.0;
  int iq = iquad_fast(d2min, fmax(d2max, floor2));
  if (iq < 0) {
    double pt[3] = {x, y, z};
    iq = iquad_exact(P, pt, 3, 1, floor2);
  }
  if (iq > 10) return phipot(P, nhat, x, y, z);
  const int n = c_qnp[iq];
  const double* bp = g_qpts + 3 * c_qoff[iq];
  const double* bw = g_qwts + c_qoff[iq];
  double pot = 0.0;
  for (int q = 0; q < n; q++) {
    double b0 = bp[3 * q], b1 = bp[3 * q + 1], b2 = bp[3 * q + 2];
    double dx = (b0 * P[0] + b1 * P[3] + b2 * P[6]) - x;
    double dy = (b0 * P[1] + b1 * P[4] + b2 * P[7]) - y;
    double dz = (b0 * P[2] + b1 * P[5] + b2 * P[8]) - z;
    pot = fma(bw[q], rsqrt_fast(fma(dz, dz, fma(dy, dy, dx * dx))), pot);
  }
  return pot * area;
}

// one warp per (cell, filament); lanes stride over the polyline points, the potential of the
// previous point comes from the neighbouring lane (or the carry of the previous 32-point block)
__global__ void elem_filament_kernel(int nc, int nfil, const double* __restrict__ cellP, const double* __restrict__ cellA,
                                     const double* __restrict__ cellE, const int* __restrict__ cell_mask,
                                     const int* __restrict__ fil_ptr, const double* __restrict__ pts,
                                     const double* __restrict__ fscale, const double* __restrict__ fradius,
                                     const int* __restrict__ fil_set, double* __restrict__ out /*[nc][nfil][3]*/,
                                     int* __restrict__ nrad_cross) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)nc * nfil) return;
  const int cell = (int)(w / nfil), f = (int)(w - (long long)cell * nfil);
  double* o = out + ((size_t)cell * nfil + f) * 3;
  if (cell_mask && cell_mask[cell]) {
    if (lane < 3) o[lane] = 0.0;
    return;
  }
  double P[9], E[9], nh[3];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    P[k] = cellP[9 * (size_t)cell + k];
    E[k] = cellE[9 * (size_t)cell + k];
  }
  tri_normal(P, nh);
  const double area = cellA[cell];
  const int p0 = fil_ptr[f], p1 = fil_ptr[f + 1];
  const double rad = fradius ? fradius[f] : -1.0;
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, carry = 0.0;
  int ncross = 0;
  for (int base = p0; base < p1; base += 32) {
    const int kk = base + lane;
    double pot = 0.0, x = 0.0, y = 0.0, z = 0.0;
    if (kk < p1) {
      x = pts[3 * (size_t)kk];
      y = pts[3 * (size_t)kk + 1];
      z = pts[3 * (size_t)kk + 2];
      double d2min;
      pot = cell_point_pot(P, area, nh, x, y, z, &d2min);
      if (rad > 0.0 && sqrt(d2min) < rad) ncross++;
    }
    double prev = __shfl_up_sync(0xffffffffu, pot, 1);
    if (lane == 0) prev = carry;
    carry = __shfl_sync(0xffffffffu, pot, 31);
    if (kk < p1 && kk > p0) {
      double cx = x - pts[3 * (size_t)kk - 3], cy = y - pts[3 * (size_t)kk - 2], cz = z - pts[3 * (size_t)kk - 1];
      double avg = (pot + prev) / 2.0;
      t0 += (E[0] * cx + E[1] * cy + E[2] * cz) * avg;
      t1 += (E[3] * cx + E[4] * cy + E[5] * cz) * avg;
      t2 += (E[6] * cx + E[7] * cy + E[8] * cz) * avg;
    }
  }
  t0 = warp_sum(t0);
  t1 = warp_sum(t1);
  t2 = warp_sum(t2);
  const double sc = fscale[f];
  if (lane == 0) {
    o[0] = sc * t0;
    o[1] = sc * t1;
    o[2] = sc * t2;
  }
  if (nrad_cross) {
    for (int ofs = 16; ofs > 0; ofs >>= 1) ncross += __shfl_xor_sync(0xffffffffu, ncross, ofs);
    if (lane == 0 && ncross) atomicAdd(&nrad_cross[fil_set[f]], ncross);  // diagnostic counter only
  }
}

// out[set][dof] (Fortran (nelems,nsets)) or out[dof][set] (Fortran (nsets,nelems)) = sum over the
// incidences of the DOF and the filaments of the set; one thread per (dof,set): no atomics.
__global__ void dof_gather_kernel(int ndof, int nsets, int nfil, const int* __restrict__ kdi, const int* __restrict__ ldi,
                                  const int* __restrict__ set_ptr, const double* __restrict__ cf /*[nc][nfil][3]*/,
                                  double* __restrict__ out, long long stride_dof, long long stride_set,
                                  const double* __restrict__ set_scale) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)ndof * nsets) return;
  const int dof = (int)(t % ndof), j = (int)(t / ndof);
  double acc = 0.0;
  for (int ii = kdi[dof]; ii < kdi[dof + 1]; ii++) {
    const int code = ldi[ii], cell = code >> 3, k = code & 3;
    double s = 0.0;
    for (int f = set_ptr[j]; f < set_ptr[j + 1]; f++) s += cf[((size_t)cell * nfil + f) * 3 + k];
    acc += (code & 4) ? -s : s;
  }
  if (set_scale) acc *= set_scale[j];
  out[dof * stride_dof + j * stride_set] = acc;
}

// one block per (row set l, column set j)
__global__ void filament_mutual_kernel(int nrow_sets, int ncol_sets, const int* __restrict__ rset_ptr,
                                       const int* __restrict__ rfil_ptr, const double* __restrict__ rpts,
                                       const double* __restrict__ rscale, const double* __restrict__ rradius,
                                       const int* __restrict__ cset_ptr, const int* __restrict__ cfil_ptr,
                                       const double* __restrict__ cpts, const double* __restrict__ cscale,
                                       const int* __restrict__ cmask, int regularize, double* __restrict__ out) {
  __shared__ double red[256];
  const int l = blockIdx.x % nrow_sets, j = blockIdx.x / nrow_sets;
  double acc = 0.0;
  if (regularize || !(cmask && cmask[j])) {
    const double sqrt_e = sqrt(exp(1.0));
    for (int i = rset_ptr[l]; i < rset_ptr[l + 1]; i++) {
      const double thick = regularize ? (rradius[i] * rradius[i]) / sqrt_e : 0.0;
      const double si = regularize ? rscale[i] : 1.0;
      for (int ii = rfil_ptr[i] + 1 + threadIdx.x; ii < rfil_ptr[i + 1]; ii += blockDim.x) {
        const double ax = rpts[3 * (size_t)ii], ay = rpts[3 * (size_t)ii + 1], az = rpts[3 * (size_t)ii + 2];
        const double bx = rpts[3 * (size_t)ii - 3], by = rpts[3 * (size_t)ii - 2], bz = rpts[3 * (size_t)ii - 1];
        const double rx = ax - bx, ry = ay - by, rz = az - bz;
        for (int k = cset_ptr[j]; k < cset_ptr[j + 1]; k++) {
          double pot_last = 0.0, tmp = 0.0;
          for (int kk = cfil_ptr[k]; kk < cfil_ptr[k + 1]; kk++) {
            const double cx = cpts[3 * (size_t)kk], cy = cpts[3 * (size_t)kk + 1], cz = cpts[3 * (size_t)kk + 2];
            double a0 = ax - cx, a1 = ay - cy, a2 = az - cz, b0 = bx - cx, b1 = by - cy, b2 = bz - cz;
            double pot = (1.0 / sqrt((a0 * a0 + a1 * a1 + a2 * a2) + thick) + 1.0 / sqrt((b0 * b0 + b1 * b1 + b2 * b2) + thick)) / 2.0;
            if (kk > cfil_ptr[k]) {
              double vx = cx - cpts[3 * (size_t)kk - 3], vy = cy - cpts[3 * (size_t)kk - 2], vz = cz - cpts[3 * (size_t)kk - 1];
              tmp += (rx * vx + ry * vy + rz * vz) * (pot + pot_last) / 2.0;
            }
            pot_last = pot;
          }
          acc += cscale[k] * si * tmp;
        }
      }
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[(size_t)j * nrow_sets + l] = red[0];
}

// midpoint Biot-Savart: out[jj*stride_c + set*stride_set + p*stride_pt] += sum (thin_wall.F90:2119-2168)
__global__ void filament_bfield_kernel(int np, int nsets, const double* __restrict__ r, const int* __restrict__ set_ptr,
                                       const int* __restrict__ fil_ptr, const double* __restrict__ pts,
                                       const double* __restrict__ fscale, const int* __restrict__ mask,
                                       double* __restrict__ out, long long stride_pt, long long stride_set,
                                       long long stride_c, double scale) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)np * nsets) return;
  const int p = (int)(t % np), j = (int)(t / np);
  if (mask && mask[j]) return;
  const double x = r[3 * (size_t)p], y = r[3 * (size_t)p + 1], z = r[3 * (size_t)p + 2];
  double e0 = 0.0, e1 = 0.0, e2 = 0.0;
  for (int k = set_ptr[j]; k < set_ptr[j + 1]; k++) {
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    for (int kk = fil_ptr[k] + 1; kk < fil_ptr[k + 1]; kk++) {
      const double* a = pts + 3 * (size_t)kk;
      const double* b = a - 3;
      double cx = a[0] - b[0], cy = a[1] - b[1], cz = a[2] - b[2];
      double vx = x - (a[0] + b[0]) / 2.0, vy = y - (a[1] + b[1]) / 2.0, vz = z - (a[2] + b[2]) / 2.0;
      double s2 = vx * vx + vy * vy + vz * vz;
      double den = s2 * sqrt(s2);
      d0 += (cy * vz - cz * vy) / den;
      d1 += (cz * vx - cx * vz) / den;
      d2 += (cx * vy - cy * vx) / den;
    }
    e0 += fscale[k] * d0;
    e1 += fscale[k] * d1;
    e2 += fscale[k] * d2;
  }
  out[0 * stride_c + j * stride_set + p * stride_pt] += e0 * scale;
  out[1 * stride_c + j * stride_set + p * stride_pt] += e1 * scale;
  out[2 * stride_c + j * stride_set + p * stride_pt] += e2 * scale;
}

// V-coil rows / columns of L (thin_wall.F90:1128-1145), scaled by 1/(4 pi)
__global__ void vcoil_fill_kernel(int nrows, const int* __restrict__ row_ids, int ns, int nv, long long nelems,
                                  const double* __restrict__ a2c /*[nv][nelems]*/, const double* __restrict__ c2c /*[nv][nv] (j*nv+i)=A(i,j)*/,
                                  double* __restrict__ out, long long ld, double scale) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nrows * (ns + nv)) return;
  const int r = (int)(t / (ns + nv)), c = (int)(t % (ns + nv));
  const int a = row_ids[r];
  if (a < ns && c < ns) return;
  double v;
  if (a < ns) v = a2c[(size_t)(c - ns) * nelems + a];           // L(a, ns+j) = Ael2coil(a,j)
  else if (c < ns) v = a2c[(size_t)(a - ns) * nelems + c];      // L(ns+i, c) = Ael2coil(c,i)
  else {
    int i = min(a, c) - ns, j = max(a, c) - ns;                 // upper triangle mirrored
    v = c2c[(size_t)j * nv + i];
  }
  out[(long long)r * ld + c] = v * scale;
}

// ---------------------------------------------------------------------------------------------
// B-field reconstruction operator: owner-computes tile = (row patch) x (64 mesh vertices)
// H[c][p] is a 3-vector with Bcontrib(vertex k of cell c) = qbasis[c][k] x H[c][p]:
//   far : H = area * sum_q w_q d_q/|d_q|^3, d_q = r_p - x_q(c)        (thin_wall.F90:2076-2088)
//   near: H = -grad phi (central differences, h = 1e-6)               (thin_wall.F90:2049-2075)
// ---------------------------------------------------------------------------------------------
constexpr int BNT = 512;
struct BelSmem {
  double g[tw::kGeomRows * tw::kCH];
  double H[3][tw::kCH * (tw::kCH + 1)];  // [comp][c][p], row stride kCH+1: the contraction reads one vertex p for the cells of 32
                                         // different DOFs at once (stride kCH would put them all in one bank)
  double vx[tw::kCH], vy[tw::kCH], vz[tw::kCH], vva[tw::kCH];
  unsigned int near_list[tw::kCH * tw::kCH];
  int dof[tw::kMaxChunkDof];
  int iptr[tw::kMaxChunkDof + 1];
  uint16_t inc[tw::kMaxChunkInc];
  unsigned long long bar;
  int near_count, item;
};

struct BelArgs {
  const tw::ChunkMeta* chunks;
  const double* geom;
  const int *chunk_dof, *inc_ptr;
  const uint16_t* inc;
  const int* patch_chunk_ptr;
  const int* row_out;      // internal dof -> local row or -1
  const double* r;         // [np][3]
  const double* va;        // [np]
  int np, nvb;             // vertices, vertex blocks
  int p0, p1;              // row patch range
  int* counter;
  double* out;             // [3][np][nrows]
  long long nrows;
  double scale;            // 1/(4 pi)
};

__device__ __forceinline__ void bel_near(const double* P, const double* nh, const double* nrm, double x, double y, double z,
                                         bool neighbor, double* Hout) {
  const double B_dx = 1.e-6;
  double pt[3] = {x, y, z}, diff[3] = {0.0, 0.0, 0.0};
  if (neighbor)
    for (int d = 0; d < 3; d++) pt[d] = pt[d] - nrm[d] * 10.0 * B_dx;
  for (int ik = 1; ik <= 2; ik++) {
    if (ik == 2)
      for (int d = 0; d < 3; d++) pt[d] = pt[d] + nrm[d] * 20.0 * B_dx;
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
      pt[jj] = pt[jj] + B_dx;
      double tmp = phipot(P, nh, pt[0], pt[1], pt[2]);
      diff[jj] = diff[jj] + tmp / (2.0 * B_dx);
      pt[jj] = pt[jj] - 2.0 * B_dx;
      tmp = phipot(P, nh, pt[0], pt[1], pt[2]);
      diff[jj] = diff[jj] - tmp / (2.0 * B_dx);
      pt[jj] = pt[jj] + B_dx;
    }
    if (!neighbor) break;
  }
  if (neighbor)
    for (int d = 0; d < 3; d++) diff[d] = diff[d] / 2.0;
  // reference: atmp = diffvec x evec = evec x (-diffvec)
  Hout[0] = -diff[0];
  Hout[1] = -diff[1];
  Hout[2] = -diff[2];
}

__global__ void __launch_bounds__(BNT, 1) bel_tile_kernel(const BelArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BelSmem& S = *reinterpret_cast<BelSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int CH = tw::kCH;
  if (tid == 0) {
    mbar_init(&S.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  const int nitems = (A.p1 - A.p0) * A.nvb;
  for (;;) {
    __syncthreads();
    if (tid == 0) S.item = atomicAdd(A.counter, 1);
    __syncthreads();
    const int item = S.item;
    if (item >= nitems) break;
    const int pa = A.p0 + item / A.nvb, vb = item % A.nvb;
    const int pbase = vb * CH, npv = min(CH, A.np - pbase);
    for (int i = tid; i < CH; i += BNT) {
      const int p = pbase + min(i, npv - 1);
      S.vx[i] = A.r[3 * (size_t)p];
      S.vy[i] = A.r[3 * (size_t)p + 1];
      S.vz[i] = A.r[3 * (size_t)p + 2];
      S.vva[i] = A.va[p];
    }
    for (int ch = A.patch_chunk_ptr[pa]; ch < A.patch_chunk_ptr[pa + 1]; ch++) {
      __syncthreads();
      const tw::ChunkMeta cm = A.chunks[ch];
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&S.bar, tw::kGeomRows * CH * 8);
        bulk_g2s(S.g, A.geom + (size_t)ch * tw::kGeomRows * CH, tw::kGeomRows * CH * 8, &S.bar);
        S.near_count = 0;
      }
      {
        const int* cdof = A.chunk_dof + cm.dof_off;
        const int* iptr = A.inc_ptr + cm.dof_off + ch;
        const uint16_t* inc = A.inc + cm.inc_off;
        for (int i = tid; i < cm.ndof; i += BNT) S.dof[i] = cdof[i];
        for (int i = tid; i <= cm.ndof; i += BNT) S.iptr[i] = iptr[i];
        const int ninc = iptr[cm.ndof];
        for (int i = tid; i < ninc; i += BNT) S.inc[i] = inc[i];
      }
      mbar_wait(&S.bar, phase);
      phase ^= 1;
      __syncthreads();
      // ---- phase 1: classify + far field; lanes = vertices, warps stride over cells
      {
        const int p = lane + 32 * (warp & 1);
        for (int c = warp >> 1; c < cm.ncell; c += BNT / 64) {
          double P[9];
#pragma unroll
          for (int k = 0; k < 9; k++) P[k] = S.g[k * CH + c];
          const double area = S.g[9 * CH + c];
          double h0 = 0.0, h1 = 0.0, h2 = 0.0;
          if (p < npv) {
            const double x = S.vx[p], y = S.vy[p], z = S.vz[p];
            double d2min = 1.e300, d2max = 0.0;
#pragma unroll
            for (int a = 0; a < 3; a++) {
              double dx = P[3 * a] - x, dy = P[3 * a + 1] - y, dz = P[3 * a + 2] - z;
              double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
              d2min = fmin(d2min, d2);
              d2max = fmax(d2max, d2);
            }
            const double floor2 = fmax(area, S.vva[p] / (3.14159265358979323846 * 3.14159265358979323846));
            int iq = iquad_fast(d2min, fmax(d2max, floor2));
            if (iq < 0) {
              double pt[3] = {x, y, z};
              iq = iquad_exact(P, pt, 3, 1, floor2);
            }
            if (iq > 10) {
              int k = atomicAdd(&S.near_count, 1);
              S.near_list[k] = (unsigned)c | ((unsigned)p << 6) | ((d2min < 0.9999999e-16 || (d2min < 1.0000001e-16 && sqrt(d2min) < 1.e-8)) ? 4096u : 0u);
            } else {
              const int n = c_qnp[iq];
              const double* bp = g_qpts + 3 * c_qoff[iq];
              const double* bw = g_qwts + c_qoff[iq];
              for (int q = 0; q < n; q++) {
                double b0 = bp[3 * q], b1 = bp[3 * q + 1], b2 = bp[3 * q + 2];
                double dx = x - (b0 * P[0] + b1 * P[3] + b2 * P[6]);
                double dy = y - (b0 * P[1] + b1 * P[4] + b2 * P[7]);
                double dz = z - (b0 * P[2] + b1 * P[5] + b2 * P[8]);
                double ri = rsqrt_fast(fma(dz, dz, fma(dy, dy, dx * dx)));
                double w3 = bw[q] * (ri * ri * ri);
                h0 = fma(w3, dx, h0);
                h1 = fma(w3, dy, h1);
                h2 = fma(w3, dz, h2);
              }
              h0 *= area;
              h1 *= area;
              h2 *= area;
            }
          }
          S.H[0][c * (CH + 1) + p] = h0;
          S.H[1][c * (CH + 1) + p] = h1;
          S.H[2][c * (CH + 1) + p] = h2;
        }
      }
      __syncthreads();
      // ---- phase 2: near field, one lane per (cell, vertex)
      for (int k = tid; k < S.near_count; k += BNT) {
        const unsigned e = S.near_list[k];
        const int c = e & 63, p = (e >> 6) & 63;
        double P[9], nh[3], nrm[3], Hn[3];
#pragma unroll
        for (int q = 0; q < 9; q++) P[q] = S.g[q * CH + c];
        tri_normal(P, nh);
        nrm[0] = S.g[22 * CH + c];
        nrm[1] = S.g[23 * CH + c];
        nrm[2] = S.g[24 * CH + c];
        bel_near(P, nh, nrm, S.vx[p], S.vy[p], S.vz[p], (e & 4096u) != 0, Hn);
        S.H[0][c * (CH + 1) + p] = Hn[0];
        S.H[1][c * (CH + 1) + p] = Hn[1];
        S.H[2][c * (CH + 1) + p] = Hn[2];
      }
      __syncthreads();
      // ---- phase 3: contraction, entry = (dof in chunk, vertex)
      {
        const int nent = cm.ndof * npv;
        for (int e = tid; e < nent; e += BNT) {
          const int p = e / cm.ndof, ia = e - p * cm.ndof;  // dof fastest: contiguous-ish writes in e
          const int row = A.row_out[S.dof[ia]];
          if (row < 0) continue;
          // (old values first: their latency overlaps the contraction)
          const size_t o = (size_t)(pbase + p) * A.nrows + row, cs = (size_t)A.np * A.nrows;
          const double o0 = __ldcg(A.out + o), o1 = __ldcg(A.out + cs + o), o2 = __ldcg(A.out + 2 * cs + o);
          double b0 = 0.0, b1 = 0.0, b2 = 0.0;
          for (int i1 = S.iptr[ia]; i1 < S.iptr[ia + 1]; i1++) {
            const unsigned w1 = S.inc[i1];
            const int c = w1 & 63, k1 = (w1 >> 6) & 3;
            const double ex = S.g[(10 + 3 * k1) * CH + c], ey = S.g[(11 + 3 * k1) * CH + c], ez = S.g[(12 + 3 * k1) * CH + c];
            const double hx = S.H[0][c * (CH + 1) + p], hy = S.H[1][c * (CH + 1) + p], hz = S.H[2][c * (CH + 1) + p];
            double c0 = ey * hz - ez * hy, c1 = ez * hx - ex * hz, c2 = ex * hy - ey * hx;
            if (w1 & 256) {
              c0 = -c0;
              c1 = -c1;
              c2 = -c2;
            }
            b0 += c0;
            b1 += c1;
            b2 += c2;
          }
          __stcg(A.out + o, o0 + b0 * A.scale);
          __stcg(A.out + cs + o, o1 + b1 * A.scale);
          __stcg(A.out + 2 * cs + o, o2 + b2 * A.scale);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// statistics of the reference loop nest: iquad histogram of the visited (not skipped) pairs
// ---------------------------------------------------------------------------------------------
__global__ void pair_stats_kernel(int nc, const double* __restrict__ cellP, const double* __restrict__ cellA,
                                  const int* __restrict__ imin, const int* __restrict__ jmax,
                                  unsigned long long* __restrict__ hist) {
  __shared__ unsigned long long h[19];
  if (threadIdx.x < 19) h[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x;
  double Pi[9];
#pragma unroll
  for (int k = 0; k < 9; k++) Pi[k] = cellP[9 * (size_t)i + k];
  const double ai = cellA[i];
  const int im = imin[i];
  unsigned int loc[19];
#pragma unroll
  for (int k = 0; k < 19; k++) loc[k] = 0;
  for (int j = threadIdx.x; j < nc; j += blockDim.x) {
    if (jmax[j] < im) continue;
    double Pj[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Pj[k] = cellP[9 * (size_t)j + k];
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
    double floor2 = fmax(ai, cellA[j]) * 2.0;
    int iq = iquad_fast(d2min, fmax(d2max, floor2));
    if (iq < 0) iq = iquad_exact(Pi, Pj, 3, 3, floor2);
#pragma unroll
    for (int k = 4; k < 19; k++) loc[k] += (iq == k);
  }
#pragma unroll
  for (int k = 4; k < 19; k++)
    if (loc[k]) atomicAdd(&h[k], (unsigned long long)loc[k]);
  __syncthreads();
  if (threadIdx.x < 19 && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace twk

// =============================================================================================
namespace tw {

#define CKO(call)                                                                                    \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return std::string(#call) + ": " + cudaGetErrorString(e_);                \
  } while (0)

namespace {
template <class T>
struct DBuf {
  T* p = nullptr;
  ~DBuf() { cudaFree(p); }
  std::string up(const std::vector<T>& h) {
    cudaFree(p);
    p = nullptr;
    CKO(cudaMalloc((void**)&p, std::max<size_t>(h.size(), 1) * sizeof(T)));
    if (!h.empty()) CKO(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return "";
  }
  std::string zeros(size_t n) {
    cudaFree(p);
    p = nullptr;
    CKO(cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)));
    CKO(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return "";
  }
};

struct CellArrays {  // plain per-cell arrays (AoS) for the O(nc * npts) kernels
  DBuf<double> P, A, E;
  std::string up(const Model& m) {
    std::vector<double> P_(9 * (size_t)m.nc);
    for (int c = 0; c < m.nc; c++)
      for (int k = 0; k < 3; k++)
        for (int d = 0; d < 3; d++) P_[9 * (size_t)c + 3 * k + d] = m.r[3 * (size_t)m.lc[3 * c + k] + d];
    std::string e;
    if (!(e = P.up(P_)).empty()) return e;
    if (!(e = A.up(m.ca)).empty()) return e;
    return E.up(m.qbasis);
  }
};

// DOF -> incidences (cell<<3 | neg<<2 | local vertex) for vertex and hole DOFs
void dof_incidence(const Model& m, std::vector<int>& kdi, std::vector<int>& ldi) {
  const int nd = m.np_active + m.nholes;
  kdi.assign(nd + 1, 0);
  for (int c = 0; c < m.nc; c++) {
    for (int k = 0; k < 3; k++)
      if (m.pmap[m.lc[3 * c + k]] > 0) kdi[m.pmap[m.lc[3 * c + k]]]++;
    for (int ii = m.kfh[c]; ii < m.kfh[c + 1]; ii++) kdi[m.np_active + std::abs(m.lfh[2 * ii])]++;
  }
  for (int i = 0; i < nd; i++) kdi[i + 1] += kdi[i];
  ldi.resize(kdi[nd]);
  std::vector<int> fill(kdi.begin(), kdi.end() - 1);
  for (int c = 0; c < m.nc; c++) {
    for (int k = 0; k < 3; k++) {
      int p = m.pmap[m.lc[3 * c + k]];
      if (p > 0) ldi[fill[p - 1]++] = (c << 3) | k;
    }
    for (int ii = m.kfh[c]; ii < m.kfh[c + 1]; ii++) {
      int h = m.lfh[2 * ii];
      ldi[fill[m.np_active + std::abs(h) - 1]++] = (c << 3) | (h < 0 ? 4 : 0) | m.lfh[2 * ii + 1];
    }
  }
}

struct DevCoils {
  DBuf<int> set_ptr, fil_ptr, mask, fil_set;
  DBuf<double> pts, scales, radius;
  int nsets = 0, nfil = 0;
  std::string up(const FlatCoils& f) {
    nsets = f.nsets();
    nfil = f.nfil();
    std::vector<int> fs(nfil);
    for (int s = 0; s < nsets; s++)
      for (int k = f.set_ptr[s]; k < f.set_ptr[s + 1]; k++) fs[k] = s;
    std::string e;
    if (!(e = set_ptr.up(f.set_ptr)).empty()) return e;
    if (!(e = fil_ptr.up(f.fil_ptr)).empty()) return e;
    if (!(e = mask.up(f.sens_mask)).empty()) return e;
    if (!(e = fil_set.up(fs)).empty()) return e;
    if (!(e = pts.up(f.pts)).empty()) return e;
    if (!(e = scales.up(f.scales)).empty()) return e;
    return radius.up(f.radius);
  }
};

std::string need_gpu() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
    cudaGetLastError();
    return "No CUDA device available (the B200 backend has no CPU fallback)";
  }
  return gpu_init_constants();
}

// element <-> filament-set coupling into `out` with the given strides (no scaling besides set_scale)
std::string elem_sets_coupling(const Model& m, const FlatCoils& fc, bool use_region_mask, double* d_out, long long stride_dof,
                               long long stride_set, const std::vector<double>* set_scale, std::vector<int>* nrad_cross) {
  CellArrays ca;
  DevCoils dc;
  std::string e;
  if (!(e = ca.up(m)).empty()) return e;
  if (!(e = dc.up(fc)).empty()) return e;
  DBuf<int> cmask, kdi_d, ldi_d, ncross;
  if (use_region_mask) {
    std::vector<int> cm(m.nc);
    for (int c = 0; c < m.nc; c++) cm[c] = m.sens_mask[m.reg[c] - 1];
    if (!(e = cmask.up(cm)).empty()) return e;
  }
  std::vector<int> kdi, ldi;
  dof_incidence(m, kdi, ldi);
  if (!(e = kdi_d.up(kdi)).empty()) return e;
  if (!(e = ldi_d.up(ldi)).empty()) return e;
  if (!(e = ncross.zeros(dc.nsets)).empty()) return e;
  DBuf<double> cf, sscale;
  if (!(e = cf.zeros((size_t)m.nc * dc.nfil * 3)).empty()) return e;
  if (set_scale && !(e = sscale.up(*set_scale)).empty()) return e;
  const long long nwarps = (long long)m.nc * dc.nfil;
  const int threads = 256;
  const long long blocks = (nwarps * 32 + threads - 1) / threads;
  twk::elem_filament_kernel<<<(unsigned)blocks, threads>>>(m.nc, dc.nfil, ca.P.p, ca.A.p, ca.E.p, use_region_mask ? cmask.p : nullptr,
                                                            dc.fil_ptr.p, dc.pts.p, dc.scales.p, nrad_cross ? dc.radius.p : nullptr,
                                                            dc.fil_set.p, cf.p, nrad_cross ? ncross.p : nullptr);
  CKO(cudaGetLastError());
  note_launch();
  const int nd = m.np_active + m.nholes;
  const long long nt = (long long)nd * dc.nsets;
  twk::dof_gather_kernel<<<(unsigned)((nt + 255) / 256), 256>>>(nd, dc.nsets, dc.nfil, kdi_d.p, ldi_d.p, dc.set_ptr.p, cf.p, d_out,
                                                                 stride_dof, stride_set, set_scale ? sscale.p : nullptr);
  CKO(cudaGetLastError());
  note_launch();
  CKO(cudaDeviceSynchronize());
  if (nrad_cross) {
    nrad_cross->resize(dc.nsets);
    if (dc.nsets) CKO(cudaMemcpy(nrad_cross->data(), ncross.p, dc.nsets * sizeof(int), cudaMemcpyDeviceToHost));
  }
  return "";
}

std::string filament_mutual(const FlatCoils& rows, const FlatCoils& cols, bool regularize, std::vector<double>& out) {
  out.assign((size_t)std::max(rows.nsets(), 1) * std::max(cols.nsets(), 1), 0.0);
  if (rows.nsets() == 0 || cols.nsets() == 0) return "";
  DevCoils dr, dcn;
  std::string e;
  if (!(e = dr.up(rows)).empty()) return e;
  if (!(e = dcn.up(cols)).empty()) return e;
  DBuf<double> d;
  if (!(e = d.zeros(out.size())).empty()) return e;
  twk::filament_mutual_kernel<<<rows.nsets() * cols.nsets(), 256>>>(rows.nsets(), cols.nsets(), dr.set_ptr.p, dr.fil_ptr.p, dr.pts.p,
                                                                    dr.scales.p, dr.radius.p, dcn.set_ptr.p, dcn.fil_ptr.p, dcn.pts.p,
                                                                    dcn.scales.p, dcn.mask.p, regularize ? 1 : 0, d.p);
  CKO(cudaGetLastError());
  note_launch();
  CKO(cudaMemcpy(out.data(), d.p, out.size() * 8, cudaMemcpyDeviceToHost));
  return "";
}
}  // namespace

// ---------------------------------------------------------------------------------------------
std::string gpu_mcoil(Model& m) {
  std::string e = need_gpu();
  if (!e.empty()) return e;
  FlatCoils all, vc;
  for (auto& s : m.vcoils) {
    all.append(s);
    vc.append(s);
  }
  for (auto& s : m.icoils) all.append(s);
  const int ntot = all.nsets(), nv = m.n_vcoils, ni = m.n_icoils;
  const size_t N = (size_t)m.nelems;
  m.Ael2coil.alloc(N * std::max(nv, 1));
  m.Ael2dr.alloc(N * std::max(ni, 1));
  m.Acoil2coil.alloc((size_t)std::max(nv, 1) * std::max(nv, 1));
  std::vector<double> tmp(N * std::max(ntot, 1), 0.0);  // Ael2coil_tmp(nelems, ntot)
  if (ntot > 0) {
    DBuf<double> d;
    if (!(e = d.zeros(tmp.size())).empty()) return e;
    std::vector<int> ncross;
    e = elem_sets_coupling(m, all, false, d.p, 1, (long long)N, nullptr, &ncross);
    if (!e.empty()) return e;
    CKO(cudaMemcpy(tmp.data(), d.p, tmp.size() * 8, cudaMemcpyDeviceToHost));
    bool warn = false;
    for (int j = 0; j < nv; j++) warn |= ncross[j] > 0;
    if (warn) {
      std::printf("WARNING: One or more elements intersect a Vcoil within its radius, which may lead to invalid inductance values.\n");
      for (int j = 0; j < nv; j++)
        if (ncross[j] > 0) std::printf("  %8d intersecting elements for coil %6d\n", ncross[j], j + 1);
    }
  }
  for (int j = 0; j < nv; j++) std::memcpy(m.Ael2coil.p + j * N, tmp.data() + j * N, N * 8);
  for (int j = 0; j < ni; j++) std::memcpy(m.Ael2dr.p + j * N, tmp.data() + (size_t)(nv + j) * N, N * 8);
  // coil <-> coil (tw_compute_Lmat_coils)
  std::vector<double> A;  // (n_v, ntot): A[j*nv + l]
  e = filament_mutual(vc, all, true, A);
  if (!e.empty()) return e;
  for (int i = 0; i < nv; i++) {
    for (int l = 0; l < nv; l++) m.Acoil2coil.p[(size_t)i * nv + l] = A[(size_t)i * nv + l];
    m.vcoils[i].Lself = A[(size_t)i * nv + i];
  }
  const int ns = m.np_active + m.nholes;
  for (int i = 0; i < nv; i++)
    for (int jj = 0; jj < ni; jj++) m.Ael2dr.p[(size_t)jj * N + ns + i] = A[(size_t)(jj + nv) * nv + i];
  for (size_t k = 0; k < N * ni; k++) m.Ael2dr.p[k] = m.Ael2dr.p[k] * kMu0 / (4.0 * kPi);
  m.have_coil_mutuals = true;
  return "";
}

std::string gpu_msensor(Model& m, const Sensors& sens) {
  std::string e = need_gpu();
  if (!e.empty()) return e;
  const int nsn = (int)sens.floops.size(), nv = m.n_vcoils, ni = m.n_icoils;
  const size_t N = (size_t)m.nelems;
  FlatCoils fs, all;
  std::vector<double> sf;
  for (auto& fl : sens.floops) {
    CoilSet cs;
    Filament f;
    f.pts = fl.pts;
    cs.coils.push_back(f);
    fs.append(cs);
    sf.push_back(fl.scale_fac);
  }
  for (auto& s : m.vcoils) all.append(s);
  for (auto& s : m.icoils) all.append(s);
  m.Ael2sen.alloc((size_t)std::max(nsn, 1) * N);   // Fortran (nsensors, nelems)
  m.Adr2sen.alloc((size_t)std::max(nsn, 1) * std::max(ni, 1));
  m.nsensors_built = nsn;
  if (nsn == 0) return "";
  {
    DBuf<double> d;
    if (!(e = d.zeros((size_t)nsn * N)).empty()) return e;
    e = elem_sets_coupling(m, fs, true, d.p, nsn, 1, &sf, nullptr);
    if (!e.empty()) return e;
    CKO(cudaMemcpy(m.Ael2sen.p, d.p, (size_t)nsn * N * 8, cudaMemcpyDeviceToHost));
  }
  std::vector<double> A;  // Acoil2sen_tmp(nsens, ntot): A[j*nsn + i]
  e = filament_mutual(fs, all, false, A);
  if (!e.empty()) return e;
  const int ntot = all.nsets();
  for (int j = 0; j < ntot; j++)
    for (int i = 0; i < nsn; i++) A[(size_t)j * nsn + i] *= sf[i];
  for (int j = 0; j < ni; j++)
    for (int i = 0; i < nsn; i++) m.Adr2sen.p[(size_t)j * nsn + i] = A[(size_t)(nv + j) * nsn + i] * kMu0 / (4.0 * kPi);
  const int ns = m.np_active + m.nholes;
  for (int i = 0; i < nv; i++)
    for (int jj = 0; jj < nsn; jj++) m.Ael2sen.p[(size_t)(ns + i) * nsn + jj] = A[(size_t)i * nsn + jj];
  for (size_t k = 0; k < (size_t)nsn * N; k++) m.Ael2sen.p[k] = m.Ael2sen.p[k] / (4.0 * kPi);
  return "";
}

std::string gpu_fill_vcoil_block(const Model& m, const std::vector<int>& row_ids, double* d_out, long long ld, cudaStream_t stream) {
  const int nv = m.n_vcoils, ns = m.np_active + m.nholes;
  if (nv == 0 || row_ids.empty()) return "";
  int* d_rows = nullptr;
  double *d_a2c = nullptr, *d_c2c = nullptr;
  CKO(cudaMallocAsync((void**)&d_rows, row_ids.size() * sizeof(int), stream));
  CKO(cudaMallocAsync((void**)&d_a2c, (size_t)nv * m.nelems * 8, stream));
  CKO(cudaMallocAsync((void**)&d_c2c, (size_t)nv * nv * 8, stream));
  CKO(cudaMemcpyAsync(d_rows, row_ids.data(), row_ids.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  CKO(cudaMemcpyAsync(d_a2c, m.Ael2coil.p, (size_t)nv * m.nelems * 8, cudaMemcpyHostToDevice, stream));
  CKO(cudaMemcpyAsync(d_c2c, m.Acoil2coil.p, (size_t)nv * nv * 8, cudaMemcpyHostToDevice, stream));
  const long long nt = (long long)row_ids.size() * (ns + nv);
  twk::vcoil_fill_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, stream>>>((int)row_ids.size(), d_rows, ns, nv, m.nelems, d_a2c, d_c2c,
                                                                            d_out, ld, 1.0 / (4.0 * kPi));
  CKO(cudaGetLastError());
  note_launch();
  CKO(cudaStreamSynchronize(stream));  // host staging buffers above are reused by the caller
  CKO(cudaFreeAsync(d_rows, stream));
  CKO(cudaFreeAsync(d_a2c, stream));
  CKO(cudaFreeAsync(d_c2c, stream));
  return "";
}

// ---------------------------------------------------------------------------------------------
std::string bel_shard_device(Model& m, int nshards, int shard, double* d_out, cudaStream_t stream) {
  std::string e = need_gpu();
  if (!e.empty()) return e;
  int device = 0;
  CKO(cudaGetDevice(&device));
  std::shared_ptr<DeviceState> ds;
  if (!(e = ensure_device(m, device, ds)).empty()) return e;
  static thread_local int attr_dev = -1;
  if (attr_dev != device) {
    CKO(cudaFuncSetAttribute(twk::bel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(twk::BelSmem)));
    attr_dev = device;
  }
  const PatchSet& ps = m.plan->ps;
  int p0, p1;
  std::vector<int> rows;
  shard_rows(m, nshards, shard, p0, p1, rows);
  const long long nrows = (long long)rows.size();
  std::vector<int> row_out(ps.ndof, -1);
  for (int i = ps.patch_dof_ptr[p0], r = 0; i < ps.patch_dof_ptr[p1]; i++, r++) row_out[i] = r;
  CKO(cudaMemsetAsync(d_out, 0, (size_t)3 * m.np * nrows * 8, stream));
  DBuf<int> d_row_out, d_counter;
  DBuf<double> d_r, d_va;
  if (!(e = d_row_out.up(row_out)).empty()) return e;
  if (!(e = d_counter.zeros(1)).empty()) return e;
  if (!(e = d_r.up(m.r)).empty()) return e;
  if (!(e = d_va.up(m.va)).empty()) return e;
  twk::BelArgs a;
  a.chunks = ds->ps.chunks;
  a.geom = ds->ps.geom;
  a.chunk_dof = ds->ps.chunk_dof;
  a.inc_ptr = ds->ps.inc_ptr;
  a.inc = ds->ps.inc;
  a.patch_chunk_ptr = ds->ps.patch_chunk_ptr;
  a.row_out = d_row_out.p;
  a.r = d_r.p;
  a.va = d_va.p;
  a.np = m.np;
  a.nvb = (m.np + kCH - 1) / kCH;
  a.p0 = p0;
  a.p1 = p1;
  a.counter = d_counter.p;
  a.out = d_out;
  a.nrows = nrows;
  a.scale = 1.0 / (4.0 * kPi);
  int nsm = 148;
  CKO(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  const long long nitems = (long long)(p1 - p0) * a.nvb;
  if (nitems > 0) {
    twk::bel_tile_kernel<<<(unsigned)std::min<long long>(nitems, nsm), twk::BNT, sizeof(twk::BelSmem), stream>>>(a);
    CKO(cudaGetLastError());
  note_launch();
  }
  // V-coil rows (last shard): filament Biot-Savart, scaled with the rest by 1/4pi
  if (shard == nshards - 1 && m.n_vcoils > 0) {
    FlatCoils vc;
    for (auto& s : m.vcoils) vc.append(s);
    DevCoils dc;
    if (!(e = dc.up(vc)).empty()) return e;
    const long long nt = (long long)m.np * vc.nsets();
    const long long off = nrows - m.n_vcoils;
    twk::filament_bfield_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, stream>>>(m.np, vc.nsets(), d_r.p, dc.set_ptr.p, dc.fil_ptr.p,
                                                                                   dc.pts.p, dc.scales.p, dc.mask.p, d_out + off, nrows, 1,
                                                                                   (long long)m.np * nrows, 1.0 / (4.0 * kPi));
    CKO(cudaGetLastError());
  note_launch();
    CKO(cudaStreamSynchronize(stream));
  }
  CKO(cudaStreamSynchronize(stream));
  return "";
}

std::string gpu_bmat(Model& m) {
  std::string e = need_gpu();
  if (!e.empty()) return e;
  const size_t N = (size_t)m.nelems, np = (size_t)m.np;
  m.Bel.alloc(3 * np * N);
  m.Bdr.alloc(3 * np * std::max(m.n_icoils, 1));
  if (!m.Bel.p) return "Host allocation of the B-field operator failed";
  if (!(e = ensure_plan(m)).empty()) return e;
  // devices: all visible ones, or only the caller's current device when the process is one rank of a one-process-per-GPU
  // job (as thincurr_Lmat does); the caller's device is current again on return
  int cur = 0;
  CKO(cudaGetDevice(&cur));
  std::vector<int> devs;
  if (std::getenv("THINCURR_B200_ONE_DEVICE") || (std::getenv("LOCAL_RANK") && !std::getenv("THINCURR_B200_NDEV"))) devs.push_back(cur);
  else
    for (int g = 0; g < visible_devices(); g++) devs.push_back(g);
  if (devs.empty()) devs.push_back(cur);
  const int ndev = std::min((int)devs.size(), std::max(1, m.plan->ps.npatch));
  for (int g = 0; g < ndev; g++) {
    CKO(cudaSetDevice(devs[g]));
    int p0, p1;
    std::vector<int> rows;
    shard_rows(m, ndev, g, p0, p1, rows);
    if (rows.empty()) continue;
    DBuf<double> d;
    CKO(cudaMalloc((void**)&d.p, 3 * np * rows.size() * 8));
    if (!(e = bel_shard_device(m, ndev, g, d.p, 0)).empty()) return e;
    // scatter local rows to Bel(e, p, comp): e fastest in memory
    std::vector<double> h(3 * np * rows.size());
    CKO(cudaMemcpy(h.data(), d.p, h.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t cp = 0; cp < 3 * np; cp++)
      for (size_t r = 0; r < rows.size(); r++) m.Bel.p[cp * N + rows[r]] = h[cp * rows.size() + r];
  }
  CKO(cudaSetDevice(cur));
  // Bdr(np, n_icoils, 3) = mu0/4pi * Biot-Savart of the I-coils
  if (m.n_icoils > 0) {
    FlatCoils ic;
    for (auto& s : m.icoils) ic.append(s);
    DevCoils dc;
    DBuf<double> d_r, d;
    if (!(e = dc.up(ic)).empty()) return e;
    if (!(e = d_r.up(m.r)).empty()) return e;
    if (!(e = d.zeros(3 * np * m.n_icoils)).empty()) return e;
    const long long nt = (long long)np * m.n_icoils;
    twk::filament_bfield_kernel<<<(unsigned)((nt + 255) / 256), 256>>>(m.np, m.n_icoils, d_r.p, dc.set_ptr.p, dc.fil_ptr.p, dc.pts.p,
                                                                        dc.scales.p, dc.mask.p, d.p, 1, (long long)np,
                                                                        (long long)np * m.n_icoils, kMu0 / (4.0 * kPi));
    CKO(cudaGetLastError());
  note_launch();
    CKO(cudaMemcpy(m.Bdr.p, d.p, 3 * np * m.n_icoils * 8, cudaMemcpyDeviceToHost));
  }
  return "";
}

// ---------------------------------------------------------------------------------------------
std::string gpu_cross_coupling(Model& m1, Model& m2, double* Mmat_host) {
  std::string e = need_gpu();
  if (!e.empty()) return e;
  int cur = 0;  // the caller's current device
  CKO(cudaGetDevice(&cur));
  std::shared_ptr<DeviceState> d1, d2;
  if (!(e = ensure_device(m1, cur, d1)).empty()) return e;
  if (!(e = ensure_device(m2, cur, d2)).empty()) return e;
  CKO(cudaSetDevice(cur));
  const PatchSet &p1 = m1.plan->ps, &p2 = m2.plan->ps;
  const size_t N1 = (size_t)m1.nelems, N2 = (size_t)m2.nelems;
  std::vector<Tile> tiles;
  build_mutual_tiles(p1, p2, tiles);
  std::vector<int> row_out(p1.ndof);
  for (int i = 0; i < p1.ndof; i++) row_out[i] = p1.dof_orig[i];  // rows written straight in reference order
  DBuf<double> d;
  if (!(e = d.zeros(N1 * N2)).empty()) return e;
  if (!(e = gpu_lmat_tiles(d1->ps, d2->ps, tiles, row_out, false, d.p, (long long)N2, 0, nullptr)).empty()) return e;
  CKO(cudaMemcpy(Mmat_host, d.p, N1 * N2 * 8, cudaMemcpyDeviceToHost));
  return "";
}

std::string gpu_pair_stats(Model& m, int64_t* hist, int64_t* visited) {
  std::string e = need_gpu();
  if (!e.empty()) return e;
  CellArrays ca;
  if (!(e = ca.up(m)).empty()) return e;
  std::vector<int> imin(m.nc), jmax(m.nc);
  for (int c = 0; c < m.nc; c++) {
    int lo = 0x7fffffff, hi = -0x7fffffff;
    for (int k = 0; k < 3; k++) {
      int p = m.pmap[m.lc[3 * c + k]];
      lo = std::min(lo, p);
      hi = std::max(hi, p);
    }
    for (int ii = m.kfh[c]; ii < m.kfh[c + 1]; ii++) {
      int h = std::abs(m.lfh[2 * ii]) + m.np_active;
      lo = std::min(lo, h);
      hi = std::max(hi, h);
    }
    imin[c] = lo;
    jmax[c] = hi;
  }
  DBuf<int> d_imin, d_jmax;
  DBuf<unsigned long long> d_hist;
  if (!(e = d_imin.up(imin)).empty()) return e;
  if (!(e = d_jmax.up(jmax)).empty()) return e;
  if (!(e = d_hist.zeros(19)).empty()) return e;
  twk::pair_stats_kernel<<<m.nc, 256>>>(m.nc, ca.P.p, ca.A.p, d_imin.p, d_jmax.p, d_hist.p);
  CKO(cudaGetLastError());
  note_launch();
  unsigned long long h[19];
  CKO(cudaMemcpy(h, d_hist.p, sizeof h, cudaMemcpyDeviceToHost));
  int64_t tot = 0;
  for (int k = 0; k < 19; k++) {
    hist[k] = (int64_t)h[k];
    tot += hist[k];
  }
  *visited = tot;
  return "";
}

double gpu_dfma_peak(int device, double* sm_clock_mhz) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  int nsm = 148, khz = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  if (sm_clock_mhz) *sm_clock_mhz = khz / 1000.0;
  const int blocks = nsm * 8, threads = 256, iters = 1 << 16;
  double* d = nullptr;
  if (cudaMalloc((void**)&d, (size_t)blocks * threads * 8) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    twk::dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1.0e-6);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = (double)blocks * threads * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
    if (rep > 0) best = std::max(best, tf);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

// ---------------------------------------------------------------------------------------------
// operator caches (Fortran unformatted sequential, thin_wall.F90:919-987,1161-1183,582-620,755-763,
// 1436-1483,1675-1684)
// ---------------------------------------------------------------------------------------------
bool lmat_cache_read(Model& m, const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::printf(" Reading element<->element self inductance matrix\n");
  int32_t want[6] = {m.nelems, m.nc, m.hash_lc(), m.hash_lc(), m.hash_r(), m.hash_r()}, got[6];
  bool ok = funf_read_record(f, got, sizeof got) && std::memcmp(want, got, sizeof got) == 0;
  if (!ok) std::printf("   Ignoring stored matrix: Model hashes do not match\n");
  const size_t N = (size_t)m.nelems;
  if (ok) {
    m.Lmat.alloc(N * N);
    for (size_t i = 0; i < N && ok; i++) ok = funf_read_record(f, m.Lmat.p + i * N + i, (N - i) * 8);  // Lmat(i,i:N)... stored in row i (symmetric)
    if (!ok) std::printf("   Error reading matrix from file\n");
  }
  std::fclose(f);
  if (ok)
    for (size_t i = 0; i < N; i++)
      for (size_t j = i + 1; j < N; j++) m.Lmat.p[j * N + i] = m.Lmat.p[i * N + j];
  return ok;
}
void lmat_cache_write(const Model& m, const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return;
  int32_t hdr[6] = {m.nelems, m.nc, m.hash_lc(), m.hash_lc(), m.hash_r(), m.hash_r()};
  funf_write_record(f, hdr, sizeof hdr);
  const size_t N = (size_t)m.nelems;
  for (size_t i = 0; i < N; i++) funf_write_record(f, m.Lmat.p + i * N + i, (N - i) * 8);
  std::fclose(f);
}
// B-field operator cache: HDF5 file with MODEL_hash i4[4] = {nelems, nc, hash(lc), hash(r)}, Bel_X|Y|Z = Fortran
// Bel(nelems,np) => C shape [np][nelems], Bdr_X|Y|Z = Fortran Bdr(np,n_icoils) => [n_icoils][np]
// (thin_wall.F90:2175-2225; oft_io.F90 reverses the dimensions on disk).  Host layout here: Bel[3][np][nelems],
// Bdr[3][n_icoils][np], i.e. one contiguous slab per dataset.
bool bmat_cache_read(Model& m, const std::string& path) {
  FILE* probe = std::fopen(path.c_str(), "rb");
  if (!probe) return false;
  std::fclose(probe);
  std::printf(" Loading B-field operator from file: %s\n", path.c_str());
  std::vector<int32_t> got;
  const int32_t want[4] = {m.nelems, m.nc, m.hash_lc(), m.hash_r()};
  if (!read_h5_dataset_i32(path, "MODEL_hash", got).empty() || got.size() != 4 || std::memcmp(got.data(), want, sizeof want) != 0) return false;
  const size_t N = (size_t)m.nelems, np = (size_t)m.np, ni = (size_t)m.n_icoils;
  m.Bel.alloc(3 * np * N, false);
  m.Bdr.alloc(3 * np * std::max<size_t>(ni, 1), false);
  if (!m.Bel.p || !m.Bdr.p) return false;
  const char* comp[3] = {"X", "Y", "Z"};
  bool ok = true;
  for (int c = 0; c < 3 && ok; c++) ok = read_h5_dataset_f64_into(path, std::string("Bel_") + comp[c], m.Bel.p + c * np * N, np * N).empty();
  for (int c = 0; c < 3 && ok; c++) ok = read_h5_dataset_f64_into(path, std::string("Bdr_") + comp[c], m.Bdr.p + c * np * ni, np * ni).empty();
  if (!ok) {
    m.Bel.release();
    m.Bdr.release();
  }
  return ok;
}
void bmat_cache_write(const Model& m, const std::string& path) {
  std::printf(" Saving B-field operator to file: %s\n", path.c_str());
  const uint64_t N = (uint64_t)m.nelems, np = (uint64_t)m.np, ni = (uint64_t)m.n_icoils;
  const int32_t hash[4] = {m.nelems, m.nc, m.hash_lc(), m.hash_r()};
  std::vector<H5Item> items;
  items.push_back({"MODEL_hash", false, {4}, hash});
  const char* comp[3] = {"X", "Y", "Z"};
  for (int c = 0; c < 3; c++) items.push_back({std::string("Bel_") + comp[c], true, {np, N}, m.Bel.p + c * np * N});
  for (int c = 0; c < 3; c++) items.push_back({std::string("Bdr_") + comp[c], true, {ni, np}, m.Bdr.p + c * np * ni});
  std::string err = write_h5_file(path, items);
  if (!err.empty()) std::printf("   %s\n", err.c_str());
}
bool mutual_cache_read(const Model& m1, const Model& m2, double* M, const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::printf(" Reading element<->element mutual inductance matrix\n");
  int32_t want[6] = {m2.nelems, m1.nelems, m1.hash_lc(), m2.hash_lc(), m1.hash_r(), m2.hash_r()}, got[6];
  bool ok = funf_read_record(f, got, sizeof got) && std::memcmp(want, got, sizeof got) == 0;
  if (!ok) std::printf("   Ignoring stored matrix: Model hashes do not match\n");
  for (size_t i = 0; i < (size_t)m1.nelems && ok; i++) ok = funf_read_record(f, M + i * m2.nelems, (size_t)m2.nelems * 8);
  std::fclose(f);
  return ok;
}
void mutual_cache_write(const Model& m1, const Model& m2, const double* M, const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return;
  int32_t hdr[6] = {m2.nelems, m1.nelems, m1.hash_lc(), m2.hash_lc(), m1.hash_r(), m2.hash_r()};
  funf_write_record(f, hdr, sizeof hdr);
  for (size_t i = 0; i < (size_t)m1.nelems; i++) funf_write_record(f, M + i * m2.nelems, (size_t)m2.nelems * 8);
  std::fclose(f);
}
bool mcoil_cache_read(Model& m, const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::printf(" Reading coil mutual matrices\n");
  int32_t want[3] = {m.nelems, m.n_vcoils, m.n_icoils}, got[3];
  bool ok = funf_read_record(f, got, sizeof got) && std::memcmp(want, got, sizeof got) == 0;
  if (!ok) std::printf("   Ignoring stored matrix: Sizes do not match\n");
  const size_t N = (size_t)m.nelems;
  if (ok) {
    m.Ael2coil.alloc(N * std::max(m.n_vcoils, 1));
    m.Ael2dr.alloc(N * std::max(m.n_icoils, 1));
    ok = funf_read_record(f, m.Ael2coil.p, N * m.n_vcoils * 8) && funf_read_record(f, m.Ael2dr.p, N * m.n_icoils * 8);
  }
  std::fclose(f);
  // like the reference, the cached path does not restore Acoil2coil (thin_wall.F90:582-620); a model
  // with V-coils therefore still needs a fresh compute_Mcoil before compute_Lmat.
  return ok;
}
void mcoil_cache_write(const Model& m, const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return;
  int32_t hdr[3] = {m.nelems, m.n_vcoils, m.n_icoils};
  funf_write_record(f, hdr, sizeof hdr);
  funf_write_record(f, m.Ael2coil.p, (size_t)m.nelems * m.n_vcoils * 8);
  funf_write_record(f, m.Ael2dr.p, (size_t)m.nelems * m.n_icoils * 8);
  std::fclose(f);
}
bool msensor_cache_read(Model& m, int nsensors, const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::printf(" Reading sensor mutual matrices\n");
  int32_t want[4] = {m.nelems, m.n_vcoils, m.n_icoils, nsensors}, got[4];
  bool ok = funf_read_record(f, got, sizeof got) && std::memcmp(want, got, sizeof got) == 0;
  if (!ok) std::printf("   Ignoring stored matrix: Sizes do not match\n");
  if (ok) {
    m.Ael2sen.alloc((size_t)std::max(nsensors, 1) * m.nelems);
    m.Adr2sen.alloc((size_t)std::max(nsensors, 1) * std::max(m.n_icoils, 1));
    ok = funf_read_record(f, m.Ael2sen.p, (size_t)nsensors * m.nelems * 8) &&
         funf_read_record(f, m.Adr2sen.p, (size_t)nsensors * m.n_icoils * 8);
  }
  std::fclose(f);
  return ok;
}
void msensor_cache_write(const Model& m, int nsensors, const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return;
  int32_t hdr[4] = {m.nelems, m.n_vcoils, m.n_icoils, nsensors};
  funf_write_record(f, hdr, sizeof hdr);
  funf_write_record(f, m.Ael2sen.p, (size_t)nsensors * m.nelems * 8);
  funf_write_record(f, m.Adr2sen.p, (size_t)nsensors * m.n_icoils * 8);
  std::fclose(f);
}

}  // namespace tw
