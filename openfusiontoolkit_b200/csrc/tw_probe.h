/* tw_probe.h -- kernel-level probes of the TEST build (libthincurr_b200_test.so, -DTW_TEST_HOOKS); not part of the
 * product library or of the public header. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
/* Kernel-level probes for the parity tests: the device functions of the operator kernels on
 * caller-given inputs.  probe_pairs: T(i,j) with cell i the analytic side when near, and the
 * selected quadrature order (thin_wall.F90:1044-1083); Pi/Pj = [n][3][3] vertices, Ai/Aj areas;
 * mode 0 = FP64 classification + far field from the vertices, mode 1 = the tile kernel's path (FP32
 * order screen with exact fallback [iquad bit 6 set when taken], far field from point tables).
 * probe_phipot: tw_compute_phipot (thin_wall.F90:1934-1985) for tri[n][3][3], pt[n][3]. */
int thincurr_b200_probe_pairs(int n, int mode, const double* Pi, const double* Ai, const double* Pj, const double* Aj,
                              double* T, int* iquad);
int thincurr_b200_probe_phipot(int n, const double* tri, const double* pt, double* out);
int thincurr_b200_probe_rsqrt(int n, const double* x, double* y);

#ifdef __cplusplus
}
#endif
