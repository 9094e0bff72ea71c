// tw_gpu.h -- host-visible interface of the CUDA side (device mirrors + kernel launchers).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "tw_host.h"
#include "tw_plan.h"

namespace tw {

// Device mirror of a PatchSet (replicated on every GPU that builds rows of this model).
struct DevicePatchSet {
  ChunkMeta* chunks = nullptr;
  double* geom = nullptr;
  int *dmin = nullptr, *dmax = nullptr, *chunk_dof = nullptr, *inc_ptr = nullptr;
  uint16_t* inc = nullptr;
  ChunkAux* aux = nullptr;   // [nchunk] index records of the L kernel
  int nchunk = 0;
  int *patch_chunk_ptr = nullptr, *dof_orig = nullptr;
  int* dof_patch = nullptr;  // [ndof] patch of every internal DOF
  size_t bytes = 0;  // host->device bytes of the last upload
  size_t cap[11] = {0};  // allocated bytes per array: a re-upload of a plan of the same size only copies (the reference-facing
                         // builds upload the model on every call; a dozen cudaFree/cudaMalloc pairs cost more than the copies)
  std::string upload_from(const PatchSet& ps);
  void release();
  ~DevicePatchSet() { release(); }
};

struct DeviceState {
  int device = 0;
  int plan_serial = -1;  // plan the mirror `ps` was uploaded from
  DevicePatchSet ps;
  double* scratch = nullptr;  // row block of the host-buffer entry points, kept between calls
  size_t scratch_bytes = 0;
  ~DeviceState() {
    if (scratch) cudaFree(scratch);
  }
};

std::string gpu_init_constants();
// number of operator kernels launched by this library so far (bench.py's gpu_launches)
void note_launch();
long long launch_count();

// d_col_map (block builds): reference column DOF id -> output column or -1; symmetrize: copy the transposed entries
// of the owned diagonal block afterwards (self builds whose tiles defer them, Tile.flags bit 3).
// Run the tile kernel: out[row_out[internal row]][ld] += (1/4pi) sum ... ; out must be zeroed.
// h_stats (optional, 8 x u64) forces a stream sync: far pairs, near T evals, 1/r evals, phipot evals.
// sb (streamed single-device build, tw_capi.cu): every tile belongs to a band of rows, the output is the whole matrix in the
// reference layout (row_out = reference id); the kernel itself runs the mirror pass of every band as soon as the band's
// tiles are done and then sets flags[b] in mapped host memory: the band's rows from its first column on and its columns
// of all later rows are final and may leave the device while the later bands are evaluated.
struct StreamBands {
  int nbands;
  const int* d_tile_band;     // device, [ntiles]: band of every tile (the queue order is the caller's)
  const int* d_ref_patch;     // device, [N]: patch of a reference DOF id
  int* d_bands;               // device, [6 nbands]: tiles per band, first / one-past-last reference id of every band's rows
                              // (2 nbands), then 3 nbands zeroed counters
  int N;
  int* flags;                 // mapped page-locked host memory, [nbands], zeroed by the caller
};
std::string gpu_lmat_tiles(const DevicePatchSet& A, const DevicePatchSet& B, const std::vector<Tile>& tiles,
                           const std::vector<int>& row_out, bool self, double* d_out, long long ld, cudaStream_t stream,
                           unsigned long long* h_stats, const int* d_col_map = nullptr, bool symmetrize = true,
                           const StreamBands* sb = nullptr);

// dst[i-i0][orig(j)] = src[j-j0][orig(i)] for internal DOFs i in [i0,i1), j in [j0,j1): the transposed block of another
// shard or band (src may be peer memory of another device); A = patch set mirror on the device that runs the kernel.
// checker: symmetric shards -- only the entries of the tiles the other shard evaluated (sym_tile_is_mine).
std::string gpu_symmetrize_cross(const DevicePatchSet& A, int i0, int i1, int j0, int j1, double* dst, const double* src, long long ld,
                                 cudaStream_t stream, bool checker = false);

// FP64 DFMA peak microbenchmark (TFLOP/s)
double gpu_dfma_peak(int device, double* sm_clock_mhz);

}  // namespace tw
