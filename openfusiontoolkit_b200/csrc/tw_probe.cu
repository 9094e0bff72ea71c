// tw_probe.cu -- kernel-level probes used by the parity tests: the device pair integral T(i,j),
// the analytic potential and the reciprocal square root, evaluated for caller-given inputs with
// exactly the device functions the operator kernels use.  (Included by tw_unity.cu.)
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/thincurr_b200.h"
#include "tw_probe.h"
#include "tw_device.cuh"
#include "tw_gpu.h"
#include "tw_ops.h"

namespace twk {

// one warp per pair; cells are staged in slot 0 of chunk-shaped shared arrays.
// mode 0: FP64 classification + far field from the vertices; mode 1: the production path of the
// tile kernel (FP32 order screen with exact fallback, far field from point tables for rules <= 16 points)
__global__ void probe_pairs_kernel(int n, int mode, const double* __restrict__ Pi /*[n][9]*/, const double* __restrict__ Ai,
                                   const double* __restrict__ Pj, const double* __restrict__ Aj, double* __restrict__ T,
                                   int* __restrict__ iquad) {
  extern __shared__ __align__(16) unsigned char probe_smem[];
  double2* tabI = reinterpret_cast<double2*>(probe_smem);
  double2* tabJ = tabI + kTabMaxN * 2 * kCH;  // (I-side stride CI <= kCH)
  double* gI = reinterpret_cast<double*>(tabJ + kTabMaxN * 2 * kCH);
  double* gJ = gI + 10 * kCH;
  double* nI = gJ + 10 * kCH;
  float* vfI = reinterpret_cast<float*>(nI + 3 * kCH);
  float* vfJ = vfI + 9 * kCH;
  const int lane = threadIdx.x;
  for (int pair = blockIdx.x; pair < n; pair += gridDim.x) {
    __syncwarp();
    if (lane < 9) {
      gI[lane * kCH] = Pi[9 * (size_t)pair + lane];
      gJ[lane * kCH] = Pj[9 * (size_t)pair + lane];
    }
    if (lane == 9) {
      gI[9 * kCH] = Ai[pair];
      gJ[9 * kCH] = Aj[pair];
    }
    __syncwarp();
    if (lane == 0) {
      double P[9], nh[3];
      for (int k = 0; k < 9; k++) P[k] = gI[k * kCH];
      tri_normal(P, nh);
      nI[0] = nh[0];
      nI[kCH] = nh[1];
      nI[2 * kCH] = nh[2];
    }
    // local frame as in the tile kernel: midpoint of two "chunk centres" (here vertex 0 of each cell,
    // displaced so that the frame is not trivially centred), X = bound on |v - o|
    // (the displacement scales with the cells like a chunk radius does: ~3.4 sqrt(area_i + area_j))
    const double dref = sqrt(gI[9 * kCH] + gJ[9 * kCH]);
    const double ox = 0.5 * (gI[0] + gJ[0]) + 1.5 * dref, oy = 0.5 * (gI[kCH] + gJ[kCH]) - 3.0 * dref,
                 oz = 0.5 * (gI[2 * kCH] + gJ[2 * kCH]) + 0.75 * dref;
    double X = 0.0;
    for (int k = 0; k < 3; k++) {
      double ax = gI[(3 * k) * kCH] - ox, ay = gI[(3 * k + 1) * kCH] - oy, az = gI[(3 * k + 2) * kCH] - oz;
      double bx = gJ[(3 * k) * kCH] - ox, by = gJ[(3 * k + 1) * kCH] - oy, bz = gJ[(3 * k + 2) * kCH] - oz;
      X = fmax(X, fmax(sqrt(ax * ax + ay * ay + az * az), sqrt(bx * bx + by * by + bz * bz)));
    }
    __syncwarp();
    int iq;
    if (mode == 0) {
      iq = classify_pair(gI, 0, gJ, 0);
    } else {
      if (lane < 9) {
        const double o = (lane % 3) == 0 ? ox : ((lane % 3) == 1 ? oy : oz);
        vfI[lane * kCH] = (float)(gI[lane * kCH] - o);
        vfJ[lane * kCH] = (float)(gJ[lane * kCH] - o);
      }
      __syncwarp();
      float pi_[9], pj_[9];
#pragma unroll
      for (int k = 0; k < 9; k++) {
        pi_[k] = vfI[k * kCH];
        pj_[k] = vfJ[k * kCH];
      }
      iq = iquad_screen(pi_, pj_, fmaxf((float)(2.0 * gI[9 * kCH]), (float)(2.0 * gJ[9 * kCH])), (float)(X * 1.21e-7));
      if (iq < 0) iq = iquad_exact_cells(gI, 0, gJ, 0) | 64;  // bit 6: the exact path was taken
    }
    const int iqv = iq & 31;
    double v;
    if (iqv > 10) {
      v = near_pair(gI, nI, 0, gJ, 0, iqv, lane, 32, 0xffffffffu);
    } else if (mode == 0 || iqv - 4 > kTabClsMax) {
      v = far_dispatch(gI, 0, gJ, 0, iqv);
    } else {
      const int np = c_qnp[iqv];
      build_table(tabI, CI, gI, 0, 1, iqv, np, ox, oy, oz, true, lane, 32);
      build_table(tabJ, kCH, gJ, 0, 1, iqv, np, ox, oy, oz, false, lane, 32);
      __syncwarp();
      v = far_tab_dispatch(tabI, tabJ, 0, 0, iqv - 4) * gI[9 * kCH] * gJ[9 * kCH];
    }
    if (lane == 0) {
      T[pair] = v;
      iquad[pair] = iq;
    }
  }
}

__global__ void probe_phipot_kernel(int n, const double* __restrict__ tri, const double* __restrict__ pt, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double P[9], nh[3];
  for (int k = 0; k < 9; k++) P[k] = tri[9 * (size_t)i + k];
  tri_normal(P, nh);
  out[i] = phipot(P, nh, pt[3 * (size_t)i], pt[3 * (size_t)i + 1], pt[3 * (size_t)i + 2]);
}

__global__ void probe_rsqrt_kernel(int n, const double* __restrict__ x, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = rsqrt_fast(x[i]);
}

}  // namespace twk

namespace {
template <class T>
struct PBuf {
  T* p = nullptr;
  ~PBuf() { cudaFree(p); }
  bool up(const T* h, size_t n) {
    if (cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return false;
    return !h || cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
  }
};
thread_local std::string g_probe_err;
}  // namespace

extern "C" {

// T(i,j) (cell i = analytic side when near) and the selected order for n cell pairs given by
// vertex coordinates [n][3][3] and areas: tco_pair_T of the oracle, thin_wall.F90:1044-1083.
int thincurr_b200_probe_pairs(int n, int mode, const double* Pi, const double* Ai, const double* Pj, const double* Aj,
                              double* T, int* iquad) {
  std::string e = tw::gpu_init_constants();
  if (!e.empty()) return 1;
  PBuf<double> dPi, dAi, dPj, dAj, dT;
  PBuf<int> dq;
  if (!dPi.up(Pi, 9 * (size_t)n) || !dAi.up(Ai, n) || !dPj.up(Pj, 9 * (size_t)n) || !dAj.up(Aj, n) || !dT.up(nullptr, n) ||
      !dq.up(nullptr, n))
    return 2;
  const int smem = 2 * twk::kTabMaxN * 2 * tw::kCH * 16 + 23 * tw::kCH * 8 + 18 * tw::kCH * 4;
  if (cudaFuncSetAttribute(twk::probe_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 4;
  twk::probe_pairs_kernel<<<std::min(n, 148 * 2), 32, smem>>>(n, mode, dPi.p, dAi.p, dPj.p, dAj.p, dT.p, dq.p);
  if (cudaMemcpy(T, dT.p, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return 3;
  if (cudaMemcpy(iquad, dq.p, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 3;
  return 0;
}

// phi_tri(pt) for n (triangle, point) pairs: tw_compute_phipot, thin_wall.F90:1934-1985
int thincurr_b200_probe_phipot(int n, const double* tri, const double* pt, double* out) {
  std::string e = tw::gpu_init_constants();
  if (!e.empty()) return 1;
  PBuf<double> dt, dp, dout;
  if (!dt.up(tri, 9 * (size_t)n) || !dp.up(pt, 3 * (size_t)n) || !dout.up(nullptr, n)) return 2;
  twk::probe_phipot_kernel<<<(n + 127) / 128, 128>>>(n, dt.p, dp.p, dout.p);
  return cudaMemcpy(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 3;
}

int thincurr_b200_probe_rsqrt(int n, const double* x, double* y) {
  PBuf<double> dx, dy;
  if (!dx.up(x, n) || !dy.up(nullptr, n)) return 2;
  twk::probe_rsqrt_kernel<<<(n + 255) / 256, 256>>>(n, dx.p, dy.p);
  return cudaMemcpy(y, dy.p, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 3;
}

}  // extern "C"
