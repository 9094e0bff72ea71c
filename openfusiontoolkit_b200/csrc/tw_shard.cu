// tw_shard.cu -- multi-device data plane of the sharded dense operators (C ABI, include/thincurr_b200.h block 3).
//
// The assembly itself needs no inter-device traffic (tw_lmat.cu).  What follows it does:
//   * exchange  -- a symmetric shard computed its diagonal block and its half of the tiles of every block it shares with
//                  another shard (a checkerboard over the patch pairs); the missing entries L[rows r][DOFs of shard s]
//                  are the transposes of entries shard s holds (thin_wall.F90:1146-1151).  The owner of the rows READS
//                  them from its peers' memory over
//                  NVLink/NVSwitch (symmetrize_cross_kernel: 32x32 transposing tiles, coalesced on both sides) -- peer
//                  pointers of the same process, or cudaIpc-mapped pointers of other ranks.
//   * gather    -- "one gather over NVLink when the full matrix is requested on one device": every shard's rows are
//                  pulled from peer memory into the reference row order of a full matrix on the calling device.
//   * export    -- rows of a shard streamed through pinned buffers into a host matrix in the reference layout or into
//                  an `Lmat.save` cache file (thin_wall.F90:1161-1171: upper-packed records, written at their file
//                  offsets so that all ranks write concurrently).
// Cross-process ordering (peer rows complete before they are read; not overwritten while being read) is the caller's:
// one stream-ordered collective (e.g. an NCCL all-reduce of one element) before and after the exchange.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/thincurr_b200.h"
#include "tw_gpu.h"
#include "tw_ops.h"

namespace twk {

// dst[row_ids[r]][0:n] = src[r][0:n]: rows of a shard (possibly peer memory) into their reference positions
__global__ void gather_rows_kernel(int nrows, const int* __restrict__ row_ids, const double* __restrict__ src, long long ld_src,
                                   double* __restrict__ dst, long long ld_dst, int n) {
  const int n2 = n >> 1;
  for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
    const double* s = src + (long long)r * ld_src;
    double* d = dst + (long long)row_ids[r] * ld_dst;
    const bool vec = ((((unsigned long long)s) | ((unsigned long long)d)) & 15ull) == 0;
    if (vec) {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
        reinterpret_cast<double2*>(d)[i] = __ldcs(reinterpret_cast<const double2*>(s) + i);
      if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) d[n - 1] = s[n - 1];
    } else {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = __ldcs(s + i);
    }
  }
}

}  // namespace twk

using namespace tw;

namespace {
int sfail(const std::string& msg) { return tw::capi_fail(msg); }
struct IntsOnDevice {
  int* d = nullptr;
  cudaStream_t s;
  IntsOnDevice(const std::vector<int>& h, cudaStream_t st) : s(st) {
    if (cudaMallocAsync((void**)&d, std::max<size_t>(h.size(), 1) * sizeof(int), st) != cudaSuccess) {
      d = nullptr;
      return;
    }
    pinned.assign(h.begin(), h.end());
    cudaMemcpyAsync(d, pinned.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);  // (pageable staging vector dies with this object)
  }
  ~IntsOnDevice() {
    if (d) cudaFreeAsync(d, s);
  }
  std::vector<int> pinned;
};
}  // namespace

extern "C" {

// ---- library-owned device memory that can be shared between ranks -----------------------------------------
int thincurr_b200_device_alloc(int64_t bytes, void** d_ptr) {
  *d_ptr = nullptr;
  if (cudaMalloc(d_ptr, (size_t)std::max<int64_t>(bytes, 8)) != cudaSuccess)
    return sfail(std::string("cudaMalloc failed: ") + cudaGetErrorString(cudaGetLastError()));
  return 0;
}
int thincurr_b200_device_free(void* d_ptr) {
  if (d_ptr && cudaFree(d_ptr) != cudaSuccess) return sfail(std::string("cudaFree failed: ") + cudaGetErrorString(cudaGetLastError()));
  return 0;
}
int thincurr_b200_ipc_export(void* d_ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, d_ptr) != cudaSuccess) return sfail(std::string("cudaIpcGetMemHandle failed: ") + cudaGetErrorString(cudaGetLastError()));
  std::memcpy(handle64, &h, 64);
  return 0;
}
int thincurr_b200_ipc_open(const unsigned char* handle64, void** d_ptr) {
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  *d_ptr = nullptr;
  if (cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
    return sfail(std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError()));
  return 0;
}
int thincurr_b200_ipc_close(void* d_ptr) {
  if (d_ptr && cudaIpcCloseMemHandle(d_ptr) != cudaSuccess) return sfail(std::string("cudaIpcCloseMemHandle failed: ") + cudaGetErrorString(cudaGetLastError()));
  return 0;
}
// same-process peers (one process driving several devices): make `peer_device` readable from the current device
int thincurr_b200_enable_peer(int peer_device) {
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess) return sfail("No CUDA device");
  if (cur == peer_device) return 0;
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, cur, peer_device) != cudaSuccess || !can) {
    cudaGetLastError();
    return sfail("Devices cannot access each other's memory");
  }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return sfail(std::string("cudaDeviceEnablePeerAccess failed: ") + cudaGetErrorString(e));
  cudaGetLastError();
  return 0;
}

// ---- exchange ---------------------------------------------------------------------------------------------
int thincurr_b200_Lmat_exchange(void* tw_ptr, int nshards, int shard, double* d_out, int64_t ld, const double* const* peer_rows,
                                void* stream_) {
  Model& m = *(Model*)tw_ptr;
  cudaStream_t stream = (cudaStream_t)stream_;
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return sfail("No CUDA device available (there is no CPU fallback)");
  std::shared_ptr<DeviceState> ds;
  std::string err = ensure_device(m, device, ds);
  if (!err.empty()) return sfail(err);
  const PatchSet& ps = m.plan->ps;
  if (nshards < 1 || shard < 0 || shard >= nshards) return sfail("Invalid shard index");
  int p0, p1;
  shard_range_sym(ps, nshards, shard, p0, p1);
  const int i0 = ps.patch_dof_ptr[p0], i1 = ps.patch_dof_ptr[p1];
  // peers in rotated order (shard+1, shard+2, ...): at any moment every rank reads from a different source, so no
  // device's memory and NVLink egress serves all readers at once
  for (int k = 1; k < nshards; k++) {
    const int s = (shard + k) % nshards;
    int q0, q1;
    shard_range_sym(ps, nshards, s, q0, q1);
    const int j0 = ps.patch_dof_ptr[q0], j1 = ps.patch_dof_ptr[q1];
    if (j1 <= j0 || i1 <= i0) continue;
    if (!peer_rows || !peer_rows[s]) return sfail("thincurr_b200_Lmat_exchange: missing row block of another shard");
    err = gpu_symmetrize_cross(ds->ps, i0, i1, j0, j1, d_out, peer_rows[s], ld, stream, true);
    if (!err.empty()) return sfail(err);
  }
  return 0;
}

// ---- gather -----------------------------------------------------------------------------------------------
int thincurr_b200_Lmat_gather(void* tw_ptr, int nshards, int sym, const double* const* shard_rows_ptr, int64_t ld_src, double* d_full,
                              int64_t ld_full, void* stream_) {
  Model& m = *(Model*)tw_ptr;
  cudaStream_t stream = (cudaStream_t)stream_;
  std::string err = ensure_plan(m);
  if (!err.empty()) return sfail(err);
  const int N = m.nelems;
  for (int s = 0; s < nshards; s++) {
    int p0, p1;
    std::vector<int> rows;
    shard_rows(m, nshards, s, p0, p1, rows, sym != 0);
    if (rows.empty()) continue;
    if (!shard_rows_ptr[s]) return sfail("thincurr_b200_Lmat_gather: missing row block");
    IntsOnDevice ids(rows, stream);
    if (!ids.d) return sfail("Device allocation failed");
    dim3 grid((unsigned)std::min(8, (N / 2 + 255) / 256 + 1), (unsigned)std::min<size_t>(rows.size(), 16384));
    twk::gather_rows_kernel<<<grid, 256, 0, stream>>>((int)rows.size(), ids.d, shard_rows_ptr[s], ld_src, d_full, ld_full, N);
    if (cudaGetLastError() != cudaSuccess) return sfail("gather_rows_kernel launch failed");
    note_launch();
  }
  return 0;
}

// ---- export -----------------------------------------------------------------------------------------------
// rows of a shard (device memory, [nrows][ld]) into a host matrix h_full[N][ld_full] in the reference layout
int thincurr_b200_rows_to_host(void* tw_ptr, int nshards, int shard, int sym, const double* d_rows, int64_t ld, double* h_full,
                               int64_t ld_full) {
  Model& m = *(Model*)tw_ptr;
  std::string err = ensure_plan(m);
  if (!err.empty()) return sfail(err);
  int p0, p1;
  std::vector<int> rows;
  shard_rows(m, nshards, shard, p0, p1, rows, sym != 0);
  const size_t N = (size_t)m.nelems;
  // pinned double buffer; runs of consecutive reference rows move as one 2-D copy
  const size_t slab_rows = std::max<size_t>(1, ((size_t)64 << 20) / (N * 8));
  double* stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaStream_t st = nullptr;
  bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 && ok; i++)
    ok = cudaMallocHost((void**)&stage[i], slab_rows * N * 8) == cudaSuccess && cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) == cudaSuccess;
  size_t pend_r0[2] = {0, 0}, pend_n[2] = {0, 0};
  auto flush = [&](int k) {
    if (!pend_n[k]) return;
    cudaEventSynchronize(ev[k]);
    if (h_full)  // (h_full == NULL: the rows are only streamed through the pinned buffers -- throughput of an export whose
                 // consumer is not host memory, e.g. a file writer that takes the staging buffers)
      for (size_t r = 0; r < pend_n[k]; r++)
        std::memcpy(h_full + (size_t)rows[pend_r0[k] + r] * ld_full, stage[k] + r * N, N * 8);
    pend_n[k] = 0;
  };
  int k = 0;
  for (size_t r0 = 0; r0 < rows.size() && ok; r0 += slab_rows) {
    const size_t n = std::min(slab_rows, rows.size() - r0);
    flush(k);
    ok = cudaMemcpy2DAsync(stage[k], N * 8, d_rows + r0 * (size_t)ld, (size_t)ld * 8, N * 8, n, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
         cudaEventRecord(ev[k], st) == cudaSuccess;
    pend_r0[k] = r0;
    pend_n[k] = n;
    k ^= 1;
  }
  flush(k);
  flush(k ^ 1);
  for (int i = 0; i < 2; i++) {
    if (stage[i]) cudaFreeHost(stage[i]);
    if (ev[i]) cudaEventDestroy(ev[i]);
  }
  if (st) cudaStreamDestroy(st);
  if (!ok) return sfail(std::string("Row export failed: ") + cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Lmat.save (self inductance, thin_wall.F90:1161-1171): record 0 = 6 x int32 [nelems, nc, hash(lc), hash(lc), hash(r), hash(r)],
// record i+1 = Lmat(i:nelems, i) = L[i][i..N-1]; gfortran framing = 4-byte length before and after each record.
static size_t lmat_save_offset(size_t N, size_t i) { return 32 + 8 * i + 8 * (i * N - (i * (i - 1)) / 2); }

int thincurr_b200_Lmat_save_begin(void* tw_ptr, const char* path) {
  Model& m = *(Model*)tw_ptr;
  const size_t N = (size_t)m.nelems;
  if ((N - 0) * 8 >= ((size_t)1 << 31) - 9) return sfail("Rows beyond 2 GiB need sub-record splitting: use the in-memory cache writer");
  FILE* f = std::fopen(path, "wb");
  if (!f) return sfail(std::string("Cannot open ") + path);
  int32_t hdr[6] = {m.nelems, m.nc, m.hash_lc(), m.hash_lc(), m.hash_r(), m.hash_r()};
  bool ok = funf_write_record(f, hdr, sizeof hdr);
  std::fclose(f);
  if (!ok) return sfail("Header write failed");
  if (truncate(path, (off_t)lmat_save_offset(N, N)) != 0) return sfail("Cannot size the cache file");
  return 0;
}

int thincurr_b200_Lmat_save_rows(void* tw_ptr, const char* path, int nshards, int shard, int sym, const double* d_rows, int64_t ld) {
  Model& m = *(Model*)tw_ptr;
  std::string err = ensure_plan(m);
  if (!err.empty()) return sfail(err);
  int p0, p1;
  std::vector<int> rows;
  shard_rows(m, nshards, shard, p0, p1, rows, sym != 0);
  const size_t N = (size_t)m.nelems;
  const int fd = open(path, O_WRONLY);
  if (fd < 0) return sfail(std::string("Cannot open ") + path);
  const size_t slab_rows = std::max<size_t>(1, ((size_t)64 << 20) / (N * 8));
  double* stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaStream_t st = nullptr;
  bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 && ok; i++)
    ok = cudaMallocHost((void**)&stage[i], slab_rows * N * 8) == cudaSuccess && cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) == cudaSuccess;
  std::vector<char> rec;
  size_t pend_r0[2] = {0, 0}, pend_n[2] = {0, 0};
  auto flush = [&](int k) {
    if (!pend_n[k] || !ok) return;
    cudaEventSynchronize(ev[k]);
    for (size_t r = 0; r < pend_n[k] && ok; r++) {
      const size_t i = (size_t)rows[pend_r0[k] + r];
      const uint32_t nb = (uint32_t)((N - i) * 8);
      rec.resize((size_t)nb + 8);
      std::memcpy(rec.data(), &nb, 4);
      std::memcpy(rec.data() + 4, stage[k] + r * N + i, nb);
      std::memcpy(rec.data() + 4 + nb, &nb, 4);
      ok = pwrite(fd, rec.data(), rec.size(), (off_t)lmat_save_offset(N, i)) == (ssize_t)rec.size();
    }
    pend_n[k] = 0;
  };
  int k = 0;
  for (size_t r0 = 0; r0 < rows.size() && ok; r0 += slab_rows) {
    const size_t n = std::min(slab_rows, rows.size() - r0);
    flush(k);
    ok = ok && cudaMemcpy2DAsync(stage[k], N * 8, d_rows + r0 * (size_t)ld, (size_t)ld * 8, N * 8, n, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
         cudaEventRecord(ev[k], st) == cudaSuccess;
    pend_r0[k] = r0;
    pend_n[k] = n;
    k ^= 1;
  }
  flush(k);
  flush(k ^ 1);
  close(fd);
  for (int i = 0; i < 2; i++) {
    if (stage[i]) cudaFreeHost(stage[i]);
    if (ev[i]) cudaEventDestroy(ev[i]);
  }
  if (st) cudaStreamDestroy(st);
  if (!ok) return sfail("Lmat.save row export failed");
  return 0;
}

}  // extern "C"
