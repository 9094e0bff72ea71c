// tw_ops.h -- operator-level entry points implemented in tw_capi.cu / tw_ops.cu.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "tw_gpu.h"

namespace tw {

int visible_devices();
// records the message returned by thincurr_b200_last_error(); returns 1
int capi_fail(const std::string& msg);
std::string ensure_plan(Model& m, int nshards = 1);
std::string ensure_device(Model& m, int device, std::shared_ptr<DeviceState>& out);
void drop_device_state(Model& m);
void shard_rows(const Model& m, int nshards, int shard, int& p0, int& p1, std::vector<int>& row_ids, bool sym = false);
std::vector<int> band_cuts(const PatchSet& ps, int p0, int p1, int nbands);
int auto_bands(const PatchSet& ps, int p0, int p1);
std::string lmat_shard_device(Model& m, int nshards, int shard, double* d_out, long long ld, cudaStream_t stream,
                              unsigned long long* stats, bool sym = false, int nbands = 1,
                              const std::function<std::string(int, int, int)>& band_done = nullptr);

// V-coil rows/columns of L from Ael2coil / Acoil2coil (thin_wall.F90:1128-1145), scaled by 1/4pi
std::string gpu_fill_vcoil_block(const Model& m, const std::vector<int>& row_ids, double* d_out, long long ld,
                                 cudaStream_t stream);
// element<->coil and coil<->coil mutuals (tw_compute_Ael2dr + tw_compute_Lmat_coils, :567-883)
std::string gpu_mcoil(Model& m);
// element->sensor and coil->sensor mutuals (tw_compute_mutuals, :1418-1686)
std::string gpu_msensor(Model& m, const Sensors& sens);
// B-field reconstruction operators (tw_compute_Bops, :1989-2169)
std::string gpu_bmat(Model& m);
std::string bel_shard_device(Model& m, int nshards, int shard, double* d_out, cudaStream_t stream);
// mutual inductance between two models (tw_compute_LmatDirect with col_model, :887-1186)
std::string gpu_cross_coupling(Model& m1, Model& m2, double* Mmat_host);
// dense apply / leading L/R eigenmodes (tw_solve.cu)
std::string gpu_rows_apply(const double* d_rows, long long ld, int nrows, int n, const double* d_x, double* d_y, cudaStream_t stream);
std::string gpu_apply_host_matrix(const double* A, size_t nrows, size_t n, double* vals);
std::string gpu_lr_eigenmodes_host(Model& m, int neigs, double* eig_vals, double* eig_vec);
std::string lr_eigs_lanczos(int n, const int* kr, const int* lc, const double* rv, int neigs, double tol, int max_dim,
                            const std::function<std::string(const double*, double*)>& apply_L, double* eig_vals, double* eig_vec,
                            int* iters_out);
// cell-list sweeps (tw_blocks.cu): HODLR dense-block builders, matrix-free apply, reduced-model products
std::string gpu_lmatblock(Model& mr, Model& mc, int nrp, const int* row_pts, int ncp, const int* col_pts, double* out, long long ld,
                          cudaStream_t stream);
std::string gpu_lmathole(Model& m, double* out, long long ld, cudaStream_t stream);
std::string gpu_bops_block(Model& m, int nrp, const int* row_pts, int ncp, const int* col_pts, int dir, double* out, long long ld,
                           cudaStream_t stream);
std::string gpu_cross_eval(Model& m1, Model& m2, int nrhs, const double* vec1, double* vec2, long long* counts);
std::string gpu_host_matrix_multi(const double* A, size_t nrows, size_t n, int nq, const double* d_X, long long ldx, double* d_Y,
                                  long long ldy);
std::string gpu_gram(const double* d_U, long long ldu, int na, const double* d_W, long long ldw, int nb, int n, double* h_G);
std::string reduce_model(Model& m, const Sensors* sens, const std::string& filename, int neigs, const double* eig_vec, bool compute_B);
// iquad histogram + visited-pair count of the reference loop nest
std::string gpu_pair_stats(Model& m, int64_t* hist, int64_t* visited);

// operator caches in the reference's on-disk formats (Fortran unformatted sequential)
bool lmat_cache_read(Model& m, const std::string& path);
void lmat_cache_write(const Model& m, const std::string& path);
bool bmat_cache_read(Model& m, const std::string& path);
void bmat_cache_write(const Model& m, const std::string& path);
bool mutual_cache_read(const Model& m1, const Model& m2, double* M, const std::string& path);
void mutual_cache_write(const Model& m1, const Model& m2, const double* M, const std::string& path);
bool mcoil_cache_read(Model& m, const std::string& path);
void mcoil_cache_write(const Model& m, const std::string& path);
bool msensor_cache_read(Model& m, int nsensors, const std::string& path);
void msensor_cache_write(const Model& m, int nsensors, const std::string& path);

}  // namespace tw
