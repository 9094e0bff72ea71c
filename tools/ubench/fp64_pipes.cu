// FP64 pipe micro-benchmarks for the far-field design decisions (B200, sm_100a):
//   dfma3     : DFMA with three distinct register operands per instruction (no operand reuse possible)
//   dfma1     : DFMA x = fma(x, a, b) with a, b shared by the K chains (operand-reuse friendly)
//   dmma      : mma.sync.m8n8k4.f64 alone (K independent accumulator tiles per warp)
//   mix       : DFMA chains and DMMA tiles interleaved in the same warps -- if the FP64 tensor path is a pipe of its own the
//               combined rate exceeds either alone
//   farpipe   : the far-field inner loop (4-op d^2, rsqrt seed, 5 refinement ops, weighted sum), K interleaved chains
//               written so that ptxas keeps them interleaved (checked in SASS)
//   farmma    : same with d^2 taken from a DMMA tile (6 FP64-pipe ops per evaluation)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <int K>
__global__ void dfma1(double* out, int iters, double a, double b) {
  double x[K];
#pragma unroll
  for (int k = 0; k < K; k++) x[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < K; k++) x[k] = fma(x[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// three distinct register operands: x[k] = fma(y[k], z[k], x[k]) with y, z per-chain registers
template <int K>
__global__ void dfma3(double* out, int iters, double a, double b) {
  double x[K], y[K], z[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    x[k] = threadIdx.x * 1e-3 + k;
    y[k] = a + 1e-9 * (k + threadIdx.x);
    z[k] = b + 1e-9 * (2 * k + threadIdx.x);
  }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < K; k++) x[k] = fma(y[k], z[k], x[k]);
#pragma unroll
    for (int k = 0; k < K; k++) y[k] = fma(z[k], x[k], y[k]);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; k++) s += x[k] + y[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int K>
__global__ void dmma(double* out, int iters, double a, double b) {
  double c0[K], c1[K];
#pragma unroll
  for (int k = 0; k < K; k++) c0[k] = c1[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < K; k++) dmma884(c0[k], c1[k], a, b, c0[k], c1[k]);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; k++) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// KF DFMA chains + KM DMMA tiles per iteration
template <int KF, int KM>
__global__ void mix(double* out, int iters, double a, double b) {
  double x[KF], c0[KM], c1[KM];
#pragma unroll
  for (int k = 0; k < KF; k++) x[k] = threadIdx.x * 1e-3 + k;
#pragma unroll
  for (int k = 0; k < KM; k++) c0[k] = c1[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < KM; k++) dmma884(c0[k], c1[k], a, b, c0[k], c1[k]);
#pragma unroll
    for (int k = 0; k < KF; k++) x[k] = fma(x[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < KF; k++) s += x[k];
#pragma unroll
  for (int k = 0; k < KM; k++) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// far-field loop, K chains, stage by stage (the row point changes every iteration through shared memory so nothing folds)
template <int K>
__global__ void farpipe(double* out, int iters, double a, double b) {
  __shared__ double4 rowpt[64];
  if (threadIdx.x < 64) rowpt[threadIdx.x] = make_double4(a + threadIdx.x * 1e-3, b - threadIdx.x * 1e-3, a * b, 40.0 + threadIdx.x);
  __syncthreads();
  double xj[K], yj[K], zj[K], sj[K], acc[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    xj[k] = threadIdx.x * 1e-3 + k;
    yj[k] = 1.0 + k + threadIdx.x * 1e-4;
    zj[k] = 2.0 - k + threadIdx.x * 1e-5;
    sj[k] = 30.0 + k + threadIdx.x * 1e-6;
    acc[k] = 0;
  }
  for (int i = 0; i < iters; i++) {
    const double4 r = rowpt[i & 63];
    double d2[K], y0[K], e[K], h[K];
#pragma unroll
    for (int k = 0; k < K; k++) d2[k] = r.w + sj[k];
#pragma unroll
    for (int k = 0; k < K; k++) d2[k] = fma(r.z, zj[k], d2[k]);
#pragma unroll
    for (int k = 0; k < K; k++) d2[k] = fma(r.y, yj[k], d2[k]);
#pragma unroll
    for (int k = 0; k < K; k++) d2[k] = fma(r.x, xj[k], d2[k]);
#pragma unroll
    for (int k = 0; k < K; k++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[k]) : "d"(d2[k]));
#pragma unroll
    for (int k = 0; k < K; k++) h[k] = d2[k] * y0[k];
#pragma unroll
    for (int k = 0; k < K; k++) e[k] = fma(-h[k], y0[k], 1.0);
#pragma unroll
    for (int k = 0; k < K; k++) h[k] = fma(0.375, e[k], 0.5);
#pragma unroll
    for (int k = 0; k < K; k++) e[k] = e[k] * y0[k];
#pragma unroll
    for (int k = 0; k < K; k++) y0[k] = fma(e[k], h[k], y0[k]);
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = fma(r.x, y0[k], acc[k]);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; k++) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// d^2 from a DMMA tile: per iteration KM tiles (2 evaluations per lane each), 6 FP64-pipe ops per evaluation
template <int KM>
__global__ void farmma(double* out, int iters, double a, double b) {
  __shared__ double rowpt[64];
  if (threadIdx.x < 64) rowpt[threadIdx.x] = a + threadIdx.x * 1e-3;
  __syncthreads();
  double bj[KM], acc[2 * KM], cinit = 40.0 + threadIdx.x * 1e-3;
#pragma unroll
  for (int k = 0; k < KM; k++) {
    bj[k] = 1.0 + k + threadIdx.x * 1e-4;
    acc[2 * k] = acc[2 * k + 1] = 0;
  }
  for (int i = 0; i < iters; i++) {
    const double ai = rowpt[(i + (threadIdx.x & 31)) & 63];
    double d2[2 * KM], y0[2 * KM], e[2 * KM], h[2 * KM];
#pragma unroll
    for (int k = 0; k < KM; k++) dmma884(d2[2 * k], d2[2 * k + 1], ai, bj[k], cinit, cinit);
#pragma unroll
    for (int k = 0; k < 2 * KM; k++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[k]) : "d"(d2[k]));
#pragma unroll
    for (int k = 0; k < 2 * KM; k++) h[k] = d2[k] * y0[k];
#pragma unroll
    for (int k = 0; k < 2 * KM; k++) e[k] = fma(-h[k], y0[k], 1.0);
#pragma unroll
    for (int k = 0; k < 2 * KM; k++) h[k] = fma(0.375, e[k], 0.5);
#pragma unroll
    for (int k = 0; k < 2 * KM; k++) e[k] = e[k] * y0[k];
#pragma unroll
    for (int k = 0; k < 2 * KM; k++) y0[k] = fma(e[k], h[k], y0[k]);
#pragma unroll
    for (int k = 0; k < 2 * KM; k++) acc[k] = fma(ai, y0[k], acc[k]);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 2 * KM; k++) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int nsm, clk;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, 148 * 1024 * 8 * 2);
  const int iters = 20000;
  printf("SMs %d clock %d kHz; all rates per clock per SM\n", nsm, clk);
  printf("%-12s %6s %4s %14s %14s %14s\n", "kernel", "warps", "K", "FP64op/clk", "DMMA fma/clk", "evals/clk");
  int warps[] = {8, 16, 32};
  for (int w : warps) {
#define RUN(NAME, TARGS, LABEL, FOPS, MFMA, EVALS)                                            \
  {                                                                                           \
    float ms = timeit([&] { NAME<TARGS><<<nsm, w * 32>>>(out, iters, 1.0000001, 1e-7); });     \
    double cyc = (ms * 1e-3) * (clk * 1e3);                                                   \
    double thr = (double)w * 32 * iters;                                                      \
    printf("%-12s %6d %4s %14.2f %14.2f %14.3f\n", #NAME, w, LABEL, thr*(FOPS) / cyc, thr*(MFMA) / cyc, thr*(EVALS) / cyc); \
  }
#define C ,
    RUN(dfma1, 8, "8", 8, 0, 0)
    RUN(dfma3, 4, "4", 8, 0, 0)
    RUN(dfma3, 8, "8", 16, 0, 0)
    RUN(dmma, 2, "2", 0, 2 * 8, 0)
    RUN(dmma, 4, "4", 0, 4 * 8, 0)
    RUN(dmma, 8, "8", 0, 8 * 8, 0)
    RUN(mix, 8 C 1, "8+1", 8, 8, 0)
    RUN(mix, 8 C 2, "8+2", 8, 16, 0)
    RUN(mix, 8 C 4, "8+4", 8, 32, 0)
    RUN(mix, 4 C 4, "4+4", 4, 32, 0)
    RUN(farpipe, 2, "2", 20, 0, 2)
    RUN(farpipe, 3, "3", 30, 0, 3)
    RUN(farpipe, 4, "4", 40, 0, 4)
    RUN(farpipe, 6, "6", 60, 0, 6)
    RUN(farmma, 1, "1", 12, 8, 2)
    RUN(farmma, 2, "2", 24, 16, 4)
    RUN(farmma, 3, "3", 36, 24, 6)
    RUN(farmma, 4, "4", 48, 32, 8)
  }
  return 0;
}
