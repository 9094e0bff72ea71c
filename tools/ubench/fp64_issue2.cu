// FP64 pipe micro-benchmark: DFMA throughput per SM as a function of resident warps and per-thread ILP,
// plus the far-field inner loop shape (4 fma d2, rsqrt seed, 5 Newton, 1 accumulate) -- what one SM can
// sustain with few warps.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void dfma(double* out, int iters, double a, double b) {
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
// far-field shape: K independent evaluations per iteration
template <int K, int SEED>
__global__ void farshape_t(double* out, int iters, double a, double b) {
  double xj[K], yj[K], zj[K], sj[K], acc[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    xj[k] = threadIdx.x * 1e-3 + k;
    yj[k] = 1.0 + k;
    zj[k] = 2.0 - k;
    sj[k] = 30.0 + k;
    acc[k] = 0;
  }
  double ax = a, ay = b, az = a + b, as = 3.0;
  for (int i = 0; i < iters; i++) {
    double d2[K], y0[K], e[K], h[K];
#pragma unroll
    for (int k = 0; k < K; k++) d2[k] = fma(ax, xj[k], fma(ay, yj[k], fma(az, zj[k], as + sj[k])));
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (SEED == 1) y0[k] = __longlong_as_double(0x5fe6eb50c7b537a9ll - (__double_as_longlong(d2[k]) >> 1));
      else asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[k]) : "d"(d2[k]));
      if (SEED == 2) { double t; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(y0[k])); y0[k] = t; }
      if (SEED == 3) { float f = __frsqrt_rn((float)d2[k]); y0[k] = (double)f; }
    }
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
    for (int k = 0; k < K; k++) acc[k] = fma(a, y0[k], acc[k]);
    ax += 1e-9;
    as += 1e-9;
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; k++) s += acc[k];
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
  int nsm;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, 148 * 1024 * 8 * 2);
  const int iters = 20000;
  printf("SMs %d clock %d kHz\n", nsm, clk);
  printf("%-10s %6s %4s %12s %12s\n", "kernel", "warps", "ILP", "FP64op/clk/SM", "frac of 64");
  int warps[] = {12, 16, 32};
  for (int w : warps) {
#define RUN(NAME, K, OPS)                                                                  \
  {                                                                                        \
    float ms = timeit([&] { NAME<K><<<nsm, w * 32>>>(out, iters, 1.0000001, 1e-7); });      \
    double ops = (double)nsm * w * 32 * iters * K * OPS;                                    \
    double per = ops / (ms * 1e-3) / (clk * 1e3) / nsm;                                     \
    printf("%-10s %6d %4d %12.2f %12.3f\n", #NAME, w, K, per, per / 64.0);                  \
  }
#define RUNF(SEED, K)                                                                      \
  {                                                                                        \
    float ms = timeit([&] { farshape_t<K, SEED><<<nsm, w * 32>>>(out, iters, 1.0000001, 1e-7); }); \
    double ev = (double)nsm * w * 32 * iters * K;                                           \
    double per = ev / (ms * 1e-3) / (clk * 1e3) / nsm;                                      \
    printf("farshape seed%d %4d warps ILP %d: %6.3f evals/clk/SM = %5.1f cycles per warp-eval per SMSP\n", SEED, w, K, per, 8.0 / per); \
  }
    RUNF(0, 4) RUNF(1, 4) RUNF(2, 4) RUNF(3, 4)
  }
  return 0;
}
