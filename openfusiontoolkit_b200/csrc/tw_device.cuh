// tw_device.cuh -- device-side math shared by the ThinCurr operator kernels (sm_100a, FP64).
//
//  * order selection   thin_wall.F90:1055-1059 (same expression at :677-681,:1535-1539,:2041-2047)
//  * analytic potential tw_compute_phipot, thin_wall.F90:1934-1985
//  * 1/r kernel with a MUFU.RSQ64H seed + one third-order correction (5 FP64-pipe ops)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace twk {

// quadrature tables (generated header, uploaded once per device)
__constant__ double c_qpts[341 * 3];
__constant__ double c_qwts[341];
__constant__ int c_qnp[19];
__constant__ int c_qoff[19];
// order thresholds on rho = dl_min/dl_max: iquad >= k  <=>  rho <= c_thr[k-5], k = 5..18.
// c_thr holds the exact decision boundaries of the reference expression evaluated in IEEE
// double with the host libm (found by bisection at start-up); c_thr2 = c_thr^2 for the
// sqrt-free fast path.
__device__ double g_qpts[341 * 3];  // same tables in global memory (lane-divergent indexing)
__device__ double g_qwts[341];
__constant__ double c_thr[14];
__constant__ double c_thr2[14];
__constant__ float c_thr2f[14];  // same, rounded to FP32 (order screening)

// ---- reciprocal square root --------------------------------------------------------------
// MUFU.RSQ64H gives ~2^-20.4 relative error on the high word; y1 = y0(1 + e/2 + 3e^2/8) with
// e = 1 - x*y0^2 leaves a truncation error ~0.31*e^3 < 2^-62, i.e. the result is good to ~1 ulp.
__device__ __forceinline__ double rsqrt_fast(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  double t = x * y0;
  double e = fma(-t, y0, 1.0);
  double p = fma(0.375, e, 0.5);
  double q = e * y0;
  return fma(q, p, y0);
}

// ---- order selection ----------------------------------------------------------------------
// exact path: bitwise the same rho as an IEEE CPU evaluation without FMA contraction
__device__ __noinline__ int iquad_exact(const double* Pi, const double* Pj, int ni, int nj, double floor2) {
  double dmin = 1.e99, dmax = __dsqrt_rn(floor2);
  for (int a = 0; a < ni; a++)
    for (int b = 0; b < nj; b++) {
      double dx = __dsub_rn(Pi[3 * a], Pj[3 * b]), dy = __dsub_rn(Pi[3 * a + 1], Pj[3 * b + 1]),
             dz = __dsub_rn(Pi[3 * a + 2], Pj[3 * b + 2]);
      double d = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      dmin = fmin(dmin, d);
      dmax = fmax(dmax, d);
    }
  if (dmin < 1.e-8) return 18;
  double rho = __ddiv_rn(dmin, dmax);
  int iq = 4;
#pragma unroll
  for (int k = 0; k < 14; k++) iq += (rho <= c_thr[k]) ? 1 : 0;
  return iq;
}

// fast path on squared distances; returns -1 when (d2min,d2max) is within the guard band of a
// threshold (caller then takes the exact path so decisions are bit-identical to the CPU).
__device__ __forceinline__ int iquad_fast(double d2min, double d2max) {
  if (d2min < 1.0000001e-16) return (d2min < 0.9999999e-16) ? 18 : -1;
  const double band = 1.e-12;
  double lo = d2max * (1.0 - band), hi = d2max * (1.0 + band);
  if (d2min > c_thr2[0] * hi) return 4;  // most pairs
  int iq = 4;
  bool amb = false;
#pragma unroll
  for (int k = 0; k < 14; k++) {
    bool in_lo = d2min <= c_thr2[k] * lo, in_hi = d2min <= c_thr2[k] * hi;
    iq += in_lo ? 1 : 0;
    amb |= (in_lo != in_hi);
  }
  return amb ? -1 : iq;
}

// ---- analytic potential of a triangle at a point -----------------------------------------
// The formula is ill-conditioned where the point lies near the extension of an edge
// (den = |r_i||e| + r_i.e cancels), so rounding differences are amplified by up to ~1e6.  To
// reproduce the reference's CPU result to 1e-10 the arithmetic is therefore done with
// non-contracted IEEE operations (no FMA) in exactly the reference's operation order
// (DOT_PRODUCT / cross_product / magnitude evaluate left to right, thin_wall.F90:1941-1983,
// oft_local.F90:314-328); sqrt and division are IEEE-rounded, log/atan2 differ by <= 1-2 ulp.
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdot(double a0, double a1, double a2, double b0, double b1, double b2) {
  return xadd(xadd(xmul(a0, b0), xmul(a1, b1)), xmul(a2, b2));
}
// quadrature point b1*P1 + b2*P2 + b3*P3 as the reference evaluates it (thin_wall.F90:1064-1066)
__device__ __forceinline__ double xquad(double b0, double b1, double b2, double p0, double p1, double p2) {
  return xadd(xadd(xmul(b0, p0), xmul(b1, p1)), xmul(b2, p2));
}

__device__ __forceinline__ double phipot(const double* P /*[3][3]*/, const double* nhat, double x, double y, double z) {
  double r[3][3], rmag[3], c[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    r[i][0] = xsub(P[3 * i], x);
    r[i][1] = xsub(P[3 * i + 1], y);
    r[i][2] = xsub(P[3 * i + 2], z);
    rmag[i] = __dsqrt_rn(xdot(r[i][0], r[i][1], r[i][2], r[i][0], r[i][1], r[i][2]));
    double dn = xdot(nhat[0], nhat[1], nhat[2], r[i][0], r[i][1], r[i][2]);
    c[i][0] = xsub(r[i][0], xmul(dn, nhat[0]));
    c[i][1] = xsub(r[i][1], xmul(dn, nhat[1]));
    c[i][2] = xsub(r[i][2], xmul(dn, nhat[2]));
  }
  double cx = xsub(xmul(r[1][1], r[2][2]), xmul(r[1][2], r[2][1]));
  double cy = xsub(xmul(r[1][2], r[2][0]), xmul(r[1][0], r[2][2]));
  double cz = xsub(xmul(r[1][0], r[2][1]), xmul(r[1][1], r[2][0]));
  double num = xdot(r[0][0], r[0][1], r[0][2], cx, cy, cz);
  double d01 = xdot(r[0][0], r[0][1], r[0][2], r[1][0], r[1][1], r[1][2]);
  double d02 = xdot(r[0][0], r[0][1], r[0][2], r[2][0], r[2][1], r[2][2]);
  double d12 = xdot(r[1][0], r[1][1], r[1][2], r[2][0], r[2][1], r[2][2]);
  double den = xadd(xadd(xadd(xmul(xmul(rmag[0], rmag[1]), rmag[2]), xmul(d01, rmag[2])), xmul(d02, rmag[1])), xmul(d12, rmag[0]));
  double omega = xmul(2.0, atan2(num, den));
  double phi = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int j = (i + 1) % 3;
    double dv0 = xsub(r[j][0], r[i][0]), dv1 = xsub(r[j][1], r[i][1]), dv2 = xsub(r[j][2], r[i][2]);
    double tmp = __dsqrt_rn(xdot(dv0, dv1, dv2, dv0, dv1, dv2));
    double n2 = xadd(xmul(rmag[j], tmp), xdot(r[j][0], r[j][1], r[j][2], dv0, dv1, dv2));
    double d2 = xadd(xmul(rmag[i], tmp), xdot(r[i][0], r[i][1], r[i][2], dv0, dv1, dv2));
    double gam = 0.0;
    if (!(fabs(d2) < 1.e-14 || tmp < 1.e-14)) gam = __ddiv_rn(log(__ddiv_rn(n2, d2)), tmp);
    double kx = xsub(xmul(c[i][1], c[j][2]), xmul(c[i][2], c[j][1]));
    double ky = xsub(xmul(c[i][2], c[j][0]), xmul(c[i][0], c[j][2]));
    double kz = xsub(xmul(c[i][0], c[j][1]), xmul(c[i][1], c[j][0]));
    phi = xadd(phi, xmul(xdot(nhat[0], nhat[1], nhat[2], kx, ky, kz), gam));
  }
  phi = xsub(phi, xmul(xdot(nhat[0], nhat[1], nhat[2], r[0][0], r[0][1], r[0][2]), omega));
  return phi;
}

__device__ __forceinline__ void tri_normal(const double* P, double* n) {
  // nhat = unit((p2-p1) x (p3-p2)), thin_wall.F90:1942-1943
  double a0 = xsub(P[3], P[0]), a1 = xsub(P[4], P[1]), a2 = xsub(P[5], P[2]);
  double b0 = xsub(P[6], P[3]), b1 = xsub(P[7], P[4]), b2 = xsub(P[8], P[5]);
  double n0 = xsub(xmul(a1, b2), xmul(a2, b1)), n1 = xsub(xmul(a2, b0), xmul(a0, b2)), n2 = xsub(xmul(a0, b1), xmul(a1, b0));
  double m = __dsqrt_rn(xdot(n0, n1, n2, n0, n1, n2));
  n[0] = __ddiv_rn(n0, m);
  n[1] = __ddiv_rn(n1, m);
  n[2] = __ddiv_rn(n2, m);
}

// ---- bulk async copy (TMA 1-D) helpers ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace twk
