// tw_solve.cu -- dense apply y = L x on the row blocks the build left in HBM, and the leading L/R eigenmodes from it.
//
// Reference: thincurr_apply_Lmat (src/python/wrappers/thincurr_f.F90:470-497, a dense mat-vec) and
// lr_eigenmodes_arpack (src/physics/thin_wall_solvers.F90:119-224: the `neigs` largest eigenvalues of
// L x = lambda R x, ARPACK mode 2 with R^-1 from a sparse solve).  Here: the mat-vec is a HBM-bound row kernel on every
// device that holds rows (no gather of the matrix, only x and y travel), the eigen solve is a Lanczos iteration in the
// R inner product with full reorthogonalisation on the host (vectors are O(N); R is sparse, ~7 entries per row, and is
// inverted by Jacobi-preconditioned CG -- the reference's own fallback when no LU package is linked, :152-157).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/thincurr_b200.h"
#include "tw_gpu.h"
#include "tw_ops.h"

namespace twk {

// y[r] = sum_j A[r][j] x[j]; one warp per row, lanes stride the row with 16-byte loads (rows are read once: streaming
// loads; x stays in L1/L2).  Rows of 8 warps per CTA.
__global__ void __launch_bounds__(256) rows_apply_kernel(const double* __restrict__ A, long long ld, int nrows, int n,
                                                         const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= nrows) return;
  const double* a = A + (long long)row * ld;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int j = 0;
  if ((((unsigned long long)a) & 15ull) == 0 && (((unsigned long long)x) & 15ull) == 0) {
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const int n2 = n >> 1;
    int i = lane;
    for (; i + 96 < n2; i += 128) {
      const double2 v0 = __ldcs(a2 + i), v1 = __ldcs(a2 + i + 32), v2 = __ldcs(a2 + i + 64), v3 = __ldcs(a2 + i + 96);
      const double2 w0 = __ldg(x2 + i), w1 = __ldg(x2 + i + 32), w2 = __ldg(x2 + i + 64), w3 = __ldg(x2 + i + 96);
      s0 = fma(v0.x, w0.x, fma(v0.y, w0.y, s0));
      s1 = fma(v1.x, w1.x, fma(v1.y, w1.y, s1));
      s2 = fma(v2.x, w2.x, fma(v2.y, w2.y, s2));
      s3 = fma(v3.x, w3.x, fma(v3.y, w3.y, s3));
    }
    for (; i < n2; i += 32) {
      const double2 v = __ldcs(a2 + i), w = __ldg(x2 + i);
      s0 = fma(v.x, w.x, fma(v.y, w.y, s0));
    }
    j = n2 * 2;
  }
  for (int i = j + lane; i < n; i += 32) s1 = fma(__ldcs(a + i), __ldg(x + i), s1);
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[row] = s;
}

}  // namespace twk

namespace tw {

#define SCK(call)                                                                       \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) return std::string(#call) + ": " + cudaGetErrorString(e_);   \
  } while (0)

std::string gpu_rows_apply(const double* d_rows, long long ld, int nrows, int n, const double* d_x, double* d_y, cudaStream_t stream) {
  if (nrows <= 0) return "";
  twk::rows_apply_kernel<<<(nrows + 7) / 8, 256, 0, stream>>>(d_rows, ld, nrows, n, d_x, d_y);
  SCK(cudaGetLastError());
  note_launch();
  return "";
}

// y = A x for a host-resident row-major matrix A[nrows][n] (vals: x on entry, y on exit; nrows == n).  The matrix is
// streamed once through two device slabs (copy engine) while the row kernel runs on the previous slab.
std::string gpu_apply_host_matrix(const double* A, size_t nrows, size_t n, double* vals) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return "No CUDA device available (the B200 backend has no CPU fallback)";
  const size_t slab = std::max<size_t>(1, std::min(nrows, ((size_t)256 << 20) / (n * 8)));
  double *d_a[2] = {nullptr, nullptr}, *d_x = nullptr, *d_y = nullptr;
  cudaStream_t sc = nullptr, sk = nullptr;
  cudaEvent_t up[2] = {nullptr, nullptr}, used[2] = {nullptr, nullptr};
  std::string err;
  auto run = [&]() -> std::string {
    SCK(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
    SCK(cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      SCK(cudaMalloc((void**)&d_a[i], slab * n * 8));
      SCK(cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming));
      SCK(cudaEventCreateWithFlags(&used[i], cudaEventDisableTiming));
    }
    SCK(cudaMalloc((void**)&d_x, n * 8));
    SCK(cudaMalloc((void**)&d_y, nrows * 8));
    SCK(cudaMemcpyAsync(d_x, vals, n * 8, cudaMemcpyHostToDevice, sk));
    int k = 0;
    for (size_t r0 = 0; r0 < nrows; r0 += slab, k ^= 1) {
      const size_t nr = std::min(slab, nrows - r0);
      SCK(cudaStreamWaitEvent(sc, used[k], 0));  // (a never-recorded event is complete)
      SCK(cudaMemcpyAsync(d_a[k], A + r0 * n, nr * n * 8, cudaMemcpyHostToDevice, sc));
      SCK(cudaEventRecord(up[k], sc));
      SCK(cudaStreamWaitEvent(sk, up[k], 0));
      std::string e = gpu_rows_apply(d_a[k], (long long)n, (int)nr, (int)n, d_x, d_y + r0, sk);
      if (!e.empty()) return e;
      SCK(cudaEventRecord(used[k], sk));
    }
    SCK(cudaMemcpyAsync(vals, d_y, nrows * 8, cudaMemcpyDeviceToHost, sk));
    SCK(cudaStreamSynchronize(sk));
    return "";
  };
  err = run();
  for (int i = 0; i < 2; i++) {
    if (d_a[i]) cudaFree(d_a[i]);
    if (up[i]) cudaEventDestroy(up[i]);
    if (used[i]) cudaEventDestroy(used[i]);
  }
  if (d_x) cudaFree(d_x);
  if (d_y) cudaFree(d_y);
  if (sc) cudaStreamDestroy(sc);
  if (sk) cudaStreamDestroy(sk);
  cudaGetLastError();
  return err;
}

// ---- host linear algebra of the Lanczos driver (all O(N) or O(m^3) with m ~ 100) --------------------------------------
namespace {

struct Csr1 {  // 1-based CSR as thincurr_Rmat returns it
  int n;
  const int *kr, *lc;
  const double* v;
  void mul(const double* x, double* y) const {
    for (int i = 0; i < n; i++) {
      double s = 0.0;
      for (int k = kr[i] - 1; k < kr[i + 1] - 1; k++) s += v[k] * x[lc[k] - 1];
      y[i] = s;
    }
  }
};

double dot(const double* a, const double* b, int n) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int i = 0;
  for (; i + 3 < n; i += 4) {
    s0 += a[i] * b[i];
    s1 += a[i + 1] * b[i + 1];
    s2 += a[i + 2] * b[i + 2];
    s3 += a[i + 3] * b[i + 3];
  }
  for (; i < n; i++) s0 += a[i] * b[i];
  return (s0 + s1) + (s2 + s3);
}

// x = R^-1 b by Jacobi-preconditioned conjugate gradients (R is symmetric positive definite)
bool pcg(const Csr1& R, const std::vector<double>& dinv, const double* b, double* x, double tol, int maxit) {
  const int n = R.n;
  std::vector<double> r(b, b + n), z(n), p(n), q(n);
  std::fill(x, x + n, 0.0);
  const double bn = std::sqrt(dot(b, b, n));
  if (bn == 0.0) return true;
  for (int i = 0; i < n; i++) z[i] = dinv[i] * r[i];
  p = z;
  double rz = dot(r.data(), z.data(), n);
  for (int it = 0; it < maxit; it++) {
    R.mul(p.data(), q.data());
    const double a = rz / dot(p.data(), q.data(), n);
    for (int i = 0; i < n; i++) {
      x[i] += a * p[i];
      r[i] -= a * q[i];
    }
    if (std::sqrt(dot(r.data(), r.data(), n)) <= tol * bn) return true;
    for (int i = 0; i < n; i++) z[i] = dinv[i] * r[i];
    const double rz1 = dot(r.data(), z.data(), n);
    const double be = rz1 / rz;
    rz = rz1;
    for (int i = 0; i < n; i++) p[i] = z[i] + be * p[i];
  }
  return false;
}

// eigen decomposition of a small symmetric matrix (cyclic Jacobi): a[m][m] -> eigenvalues w, eigenvectors in the columns of v
void jacobi_eig(std::vector<double>& a, int m, std::vector<double>& w, std::vector<double>& v) {
  v.assign((size_t)m * m, 0.0);
  for (int i = 0; i < m; i++) v[(size_t)i * m + i] = 1.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < m; i++)
      for (int j = 0; j < m; j++) (i == j ? dg : off) += a[(size_t)i * m + j] * a[(size_t)i * m + j];
    if (off <= 1e-30 * dg) break;
    for (int p = 0; p < m - 1; p++)
      for (int q = p + 1; q < m; q++) {
        const double apq = a[(size_t)p * m + q];
        if (apq == 0.0) continue;
        const double th = (a[(size_t)q * m + q] - a[(size_t)p * m + p]) / (2.0 * apq);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < m; k++) {
          const double akp = a[(size_t)k * m + p], akq = a[(size_t)k * m + q];
          a[(size_t)k * m + p] = c * akp - s * akq;
          a[(size_t)k * m + q] = s * akp + c * akq;
        }
        for (int k = 0; k < m; k++) {
          const double apk = a[(size_t)p * m + k], aqk = a[(size_t)q * m + k];
          a[(size_t)p * m + k] = c * apk - s * aqk;
          a[(size_t)q * m + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < m; k++) {
          const double vkp = v[(size_t)k * m + p], vkq = v[(size_t)k * m + q];
          v[(size_t)k * m + p] = c * vkp - s * vkq;
          v[(size_t)k * m + q] = s * vkp + c * vkq;
        }
      }
  }
  w.resize(m);
  for (int i = 0; i < m; i++) w[i] = a[(size_t)i * m + i];
}

}  // namespace

// `neigs` largest eigenvalues of L x = lambda R x.  apply_L(x, y): y = L x (device-resident L).  eig_vec[neigs][n].
// Lanczos on A = R^-1 L, self-adjoint in <u,v>_R = u.R v: with R-orthonormal V the projected matrix is H = V^T L V.
std::string lr_eigs_lanczos(int n, const int* kr, const int* lc, const double* rv, int neigs, double tol, int max_dim,
                            const std::function<std::string(const double*, double*)>& apply_L, double* eig_vals, double* eig_vec,
                            int* iters_out) {
  if (neigs < 1 || neigs > n) return "Invalid number of eigenvalues";
  Csr1 R{n, kr, lc, rv};
  std::vector<double> dinv(n, 1.0);
  for (int i = 0; i < n; i++)
    for (int k = kr[i] - 1; k < kr[i + 1] - 1; k++)
      if (lc[k] - 1 == i && rv[k] > 0.0) dinv[i] = 1.0 / rv[k];
  max_dim = std::min(n, std::max(max_dim, 2 * neigs + 8));
  std::vector<std::vector<double>> V, LV;  // R-orthonormal basis and L V
  std::vector<double> H((size_t)max_dim * max_dim, 0.0);
  std::vector<double> w(n), t(n), u(n);
  // deterministic start vector (smooth + a fixed pseudo-random part so that no mode is missed by symmetry)
  unsigned long long s = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < n; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    w[i] = 1.0 + 0.5 * ((double)(s >> 11) / 9007199254740992.0 - 0.5);
  }
  R.mul(w.data(), t.data());
  double nrm = std::sqrt(dot(w.data(), t.data(), n));
  for (int i = 0; i < n; i++) w[i] /= nrm;
  V.push_back(w);
  std::vector<double> theta, S;
  int m = 0;
  bool converged = false;
  for (int j = 0; j < max_dim; j++) {
    std::string e = apply_L(V[j].data(), u.data());
    if (!e.empty()) return e;
    LV.push_back(u);
    for (int i = 0; i <= j; i++) {
      const double h = dot(V[i].data(), u.data(), n);
      H[(size_t)i * max_dim + j] = h;
      H[(size_t)j * max_dim + i] = h;
    }
    m = j + 1;
    if (!pcg(R, dinv, u.data(), w.data(), 1e-14, 20 * n + 1000)) return "Resistance-matrix solve did not converge (Lanczos)";
    // full reorthogonalisation in the R inner product, twice
    for (int pass = 0; pass < 2; pass++) {
      R.mul(w.data(), t.data());
      for (int i = 0; i <= j; i++) {
        const double h = dot(V[i].data(), t.data(), n);
        for (int k = 0; k < n; k++) w[k] -= h * V[i][k];
      }
    }
    R.mul(w.data(), t.data());
    const double beta = std::sqrt(std::max(0.0, dot(w.data(), t.data(), n)));
    const bool check = (m >= neigs) && (m % 4 == 0 || m == max_dim || beta < 1e-300 || m == n);
    if (check) {
      std::vector<double> Hm((size_t)m * m);
      for (int a = 0; a < m; a++)
        for (int b = 0; b < m; b++) Hm[(size_t)a * m + b] = H[(size_t)a * max_dim + b];
      jacobi_eig(Hm, m, theta, S);
      // residual of Ritz pair i: beta * |last component of s_i|
      std::vector<int> ord(m);
      for (int i = 0; i < m; i++) ord[i] = i;
      std::sort(ord.begin(), ord.end(), [&](int a, int b) { return std::fabs(theta[a]) > std::fabs(theta[b]); });
      bool ok = true;
      for (int k = 0; k < neigs; k++) {
        const int i = ord[k];
        if (beta * std::fabs(S[(size_t)(m - 1) * m + i]) > tol * std::fabs(theta[i])) ok = false;
      }
      if (ok || m == n || beta < 1e-300) {
        for (int k = 0; k < neigs; k++) {
          const int i = ord[k];
          eig_vals[k] = theta[i];
          double* x = eig_vec + (size_t)k * n;
          std::fill(x, x + n, 0.0);
          for (int a = 0; a < m; a++) {
            const double c = S[(size_t)a * m + i];
            for (int q = 0; q < n; q++) x[q] += c * V[a][q];
          }
        }
        converged = true;
        break;
      }
    }
    if (j + 1 < max_dim) {
      for (int k = 0; k < n; k++) w[k] /= beta;
      V.push_back(w);
    }
  }
  if (iters_out) *iters_out = m;
  if (!converged) return "Lanczos iteration did not converge within the basis limit";
  return "";
}

// host-resident L (library-owned buffer of thincurr_Lmat): rows are uploaded once to the devices in use and stay
// there for the iteration
std::string gpu_lr_eigenmodes_host(Model& m, int neigs, double* eig_vals, double* eig_vec) {
  const size_t N = (size_t)m.nelems;
  int ndev = std::max(1, std::getenv("LOCAL_RANK") ? 1 : visible_devices());
  if (visible_devices() < 1) return "No CUDA device available (the B200 backend has no CPU fallback)";
  int cur = 0;
  cudaGetDevice(&cur);
  ndev = (int)std::min<size_t>(ndev, std::max<size_t>(1, N / 1024));
  struct Slab {
    int dev;
    size_t r0, nr;
    double *A = nullptr, *x = nullptr, *y = nullptr;
    cudaStream_t s = nullptr;
  };
  std::vector<Slab> slabs(ndev);
  std::string err;
  auto cleanup = [&]() {
    for (auto& sl : slabs) {
      cudaSetDevice(sl.dev);
      if (sl.A) cudaFree(sl.A);
      if (sl.x) cudaFree(sl.x);
      if (sl.y) cudaFree(sl.y);
      if (sl.s) cudaStreamDestroy(sl.s);
    }
    cudaSetDevice(cur);
    cudaGetLastError();
  };
  for (int g = 0; g < ndev && err.empty(); g++) {
    Slab& sl = slabs[g];
    sl.dev = ndev == 1 ? cur : g;
    sl.r0 = N * g / ndev;
    sl.nr = N * (g + 1) / ndev - sl.r0;
    if (cudaSetDevice(sl.dev) != cudaSuccess || cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&sl.A, std::max<size_t>(sl.nr * N * 8, 8)) != cudaSuccess || cudaMalloc((void**)&sl.x, N * 8) != cudaSuccess ||
        cudaMalloc((void**)&sl.y, std::max<size_t>(sl.nr, 1) * 8) != cudaSuccess ||
        cudaMemcpyAsync(sl.A, m.Lmat.p + sl.r0 * N, sl.nr * N * 8, cudaMemcpyHostToDevice, sl.s) != cudaSuccess)
      err = std::string("Device allocation / upload of the inductance matrix failed: ") + cudaGetErrorString(cudaGetLastError());
  }
  auto apply = [&](const double* x, double* y) -> std::string {
    for (auto& sl : slabs) {
      cudaSetDevice(sl.dev);
      SCK(cudaMemcpyAsync(sl.x, x, N * 8, cudaMemcpyHostToDevice, sl.s));
      std::string e = gpu_rows_apply(sl.A, (long long)N, (int)sl.nr, (int)N, sl.x, sl.y, sl.s);
      if (!e.empty()) return e;
      SCK(cudaMemcpyAsync(y + sl.r0, sl.y, sl.nr * 8, cudaMemcpyDeviceToHost, sl.s));
    }
    for (auto& sl : slabs) {
      cudaSetDevice(sl.dev);
      SCK(cudaStreamSynchronize(sl.s));
    }
    return "";
  };
  int iters = 0;
  if (err.empty())
    err = lr_eigs_lanczos((int)N, m.R_kr.data(), m.R_lc.data(), m.R_val.data(), neigs, 1e-10, 400, apply, eig_vals, eig_vec, &iters);
  cleanup();
  if (err.empty() && m.verbose) {
    std::printf("\n Starting eigenvalue solve (Lanczos, %d device%s, %d mat-vecs)\n   Eigenvalues\n", ndev, ndev > 1 ? "s" : "", iters);
    for (int i = 0; i < std::min(neigs, 5); i++) std::printf("     %14.6E\n", eig_vals[i]);
    std::fflush(stdout);
  }
  return err;
}

}  // namespace tw

extern "C" {

int thincurr_b200_rows_apply(const double* d_rows, int64_t ld, int nrows, int n, const double* d_x, double* d_y, void* stream) {
  std::string e = tw::gpu_rows_apply(d_rows, (long long)ld, nrows, n, d_x, d_y, (cudaStream_t)stream);
  return e.empty() ? 0 : tw::capi_fail(e);
}

int thincurr_b200_lr_eigs(void* tw_ptr, int neigs, double tol, int max_dim, thincurr_b200_apply_fn apply, void* user, double* eig_vals,
                          double* eig_vec, int* n_applies) {
  tw::Model& m = *(tw::Model*)tw_ptr;
  if (m.R_kr.empty()) return tw::capi_fail("Resistance matrix required, but not computed");
  if (!apply) return tw::capi_fail("thincurr_b200_lr_eigs: no apply callback");
  auto f = [&](const double* x, double* y) -> std::string {
    return apply(user, x, y) == 0 ? std::string() : std::string("apply callback failed");
  };
  std::string e = tw::lr_eigs_lanczos(m.nelems, m.R_kr.data(), m.R_lc.data(), m.R_val.data(), neigs, tol > 0 ? tol : 1e-10,
                                      max_dim > 0 ? max_dim : 400, f, eig_vals, eig_vec, n_applies);
  return e.empty() ? 0 : tw::capi_fail(e);
}

}  // extern "C"
