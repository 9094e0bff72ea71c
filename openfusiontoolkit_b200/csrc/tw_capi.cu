// tw_capi.cu -- C ABI of libthincurr_b200.so (see include/thincurr_b200.h).
//
// Block 1 mirrors the BIND(C) wrappers of the reference (src/python/wrappers/thincurr_f.F90,
// oft_base_f.F90) for the operator-build path; block 2 is the sharded device interface.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/thincurr_b200.h"
#include "tw_gpu.h"
#include "tw_ops.h"

using namespace tw;

namespace {
thread_local std::string g_last_error;
int g_debug = 0;

void set_err(char* error_str, const std::string& msg) {
  if (!error_str) return;
  std::snprintf(error_str, THINCURR_ERROR_SLEN, "%s", msg.c_str());
}
int fail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}
std::string cstr(const char* s) {
  if (!s) return "";
  std::string out(s, strnlen(s, THINCURR_PATH_SLEN));
  while (!out.empty() && (out.back() == ' ' || out.back() == '\n')) out.pop_back();
  return out;
}
std::string time_to_string(double s) {
  // "  Time = " lines of the reference print hh:mm:ss style strings (oft_local.F90 time_to_string)
  int hours = (int)(s / 3600.0), minutes = (int)((s - hours * 3600.0) / 60.0);
  double seconds = s - hours * 3600.0 - minutes * 60.0;
  char buf[64];
  if (hours > 0) std::snprintf(buf, sizeof buf, "%dh %dm %.0fs", hours, minutes, seconds);
  else if (minutes > 0) std::snprintf(buf, sizeof buf, "%dm %.0fs", minutes, seconds);
  else std::snprintf(buf, sizeof buf, "%.3fs", seconds);
  return buf;
}
struct XmlDoc {
  std::unique_ptr<XmlNode> root;
};
}  // namespace

// ---------------------------------------------------------------------------------------------
// helpers shared with tw_ops.cu
// ---------------------------------------------------------------------------------------------
namespace tw {

int capi_fail(const std::string& msg) { return fail(msg); }

void HostBuf::alloc(size_t count, bool zero) {
  if (p && n == count) {  // reuse (rebuilds of the same operator keep the Python view valid)
    if (zero) std::memset(p, 0, count * sizeof(double));
    return;
  }
  release();
  n = count;
  if (count == 0) return;
  size_t bytes = count * sizeof(double);
  // page-locked memory (direct copy-engine target) up to THINCURR_B200_PIN_MAX_GB (default 192 GiB); beyond that, or if
  // the host refuses to lock that much, a pageable buffer served by the staging threads of HostCopier
  size_t pin_max = (size_t)192 << 30;
  if (const char* e = std::getenv("THINCURR_B200_PIN_MAX_GB")) pin_max = (size_t)std::max(0, std::atoi(e)) << 30;
  if (bytes <= pin_max) {
    if (cudaMallocHost((void**)&p, bytes) == cudaSuccess) {
      pinned = true;
      if (zero) std::memset(p, 0, bytes);
      return;
    }
    cudaGetLastError();
  }
  p = (double*)std::calloc(count, sizeof(double));
  pinned = false;
}
void HostBuf::release() {
  if (p) {
    if (pinned) cudaFreeHost(p);
    else std::free(p);
  }
  p = nullptr;
  n = 0;
}

int visible_devices() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (const char* e = std::getenv("THINCURR_B200_NDEV")) n = std::max(1, std::min(n, std::atoi(e)));
  return n;
}

// nshards: how many devices the first build that needs the plan spreads the rows over (sizes the patches so that every
// device still gets enough tiles); the plan is kept for the lifetime of the model whatever later calls ask for
static int g_plan_serial = 0;
std::string ensure_plan(Model& m, int nshards) {
  if (m.plan) return "";
  auto pl = std::make_shared<Plan>();
  int P = 0;
  if (const char* e = std::getenv("THINCURR_B200_PATCH")) P = std::atoi(e);
  std::string err = build_patches(m, P, pl->ps, std::max(1, nshards));
  if (!err.empty()) return err;
  pl->nshards_hint = std::max(1, nshards);
  pl->serial = ++g_plan_serial;
  m.plan = pl;
  return "";
}
// Plan sized for `nshards` pieces of work even if another one exists (builds whose rows leave the device band by band
// need many more tiles than a device-resident build); device mirrors follow on their next use.
std::string replan(Model& m, int nshards) {
  if (m.plan && m.plan->nshards_hint == std::max(1, nshards)) return "";
  m.plan.reset();
  return ensure_plan(m, nshards);
}

// Banded plan of the streamed single-device build (lmat_stream_host): the vertex DOFs are cut into B equal ranges of
// reference ids, every range gets patches of its own.  Band b evaluates its rows against itself and all later bands; with
// uniform cost per entry a band holding the rows [F0,F1) (fractions of N) carries the share (1-F0)^2 - (1-F1)^2 of the work
// AND of the bytes that can leave the device when it is done (its rows from its own columns on, plus its columns of all
// later rows), so the copies keep pace with the evaluation; what stays exposed is the copy of the last band, 1/B^2 of the
// matrix.  B <= 10, and few enough bands that a range of reference ids (on a structured mesh: a strip of grid lines) is
// not narrower than a patch: sqrt(nv / (1.1 P)).  Empty result: no streaming (small or V-coil models).
static std::vector<int> stream_ref_cuts(const Model& m, int P, int ndev) {
  const int nv = m.np_active;
  int B = 0;
  if (const char* e = std::getenv("THINCURR_B200_STREAM_BANDS")) B = std::atoi(e);
  else if (m.nelems >= 12000 && m.n_vcoils == 0) B = std::min(10, (int)std::sqrt(nv / (1.125 * std::max(P, 1))));
  // several devices: the bands are dealt out to the devices (a band needs nothing from any other band), at least four
  // per device so that their unequal costs balance (narrower ranges: more halo, paid for by the second link)
  if (ndev > 1 && B >= 2) B = std::max(B, 4 * ndev);
  if (B < 2 || nv < 64 * B || m.n_vcoils > 0) return {};
  B = std::min(B, 32);
  std::vector<int> cuts{0};
  for (int b = 1; b < B; b++) {
    const int c = std::min((int)std::lround((double)nv * b / B / 32.0) * 32, nv - 32);
    if (c > cuts.back()) cuts.push_back(c);
  }
  cuts.push_back(nv);
  return cuts.size() >= 3 ? cuts : std::vector<int>{};
}
// (re)plan for the streamed build; false: this model is built the ordinary way
static bool ensure_banded_plan(Model& m, std::string& err, int ndev = 1) {
  err.clear();
  int P = 0;
  if (const char* e = std::getenv("THINCURR_B200_PATCH")) P = std::atoi(e);
  if (m.plan && !m.plan->band_ref_ptr.empty() && (P <= 0 || m.plan->patch_size == P) && m.plan->band_ndev == ndev) return true;
  if (m.no_stream_plan) return false;
  // (one device, 100k-vertex vessel: 600 -> 1.98 s, 850 -> 1.93 s, 1200 -> 2.17 s end to end)
  if (P <= 0) P = auto_patch_size(m.np_active, 4 * ndev);
  const std::vector<int> cuts = stream_ref_cuts(m, P, ndev);
  if (cuts.empty()) return false;
  auto pl = std::make_shared<Plan>();
  err = build_patches(m, P, pl->ps, 1, &cuts, &pl->band_patch_ptr);
  if (!err.empty()) return false;
  {
    // A range of reference ids must be a compact piece of the surface for its patches to be compact: with a numbering
    // without locality (a permuted mesh) the one-ring halos of the patches multiply the cells, hence the pair work.
    // Compact patches of >= 300 DOFs carry 1.1-1.3 x nc cells in total.
    double cells = 0.0;
    for (int p = 0; p < pl->ps.nvert_patch; p++) cells += pl->ps.patch_ncell[p];
    if (cells > 1.6 * std::max(m.nc, 1) && !std::getenv("THINCURR_B200_STREAM_BANDS")) {
      m.no_stream_plan = true;
      return false;
    }
  }
  pl->band_ref_ptr = cuts;
  pl->band_ref_ptr.back() = m.np_active + m.nholes;  // the hole DOFs (reference ids after the vertices) and their patches
  pl->band_patch_ptr.back() = pl->ps.npatch;         // belong to the last band
  pl->nshards_hint = ndev;
  pl->band_ndev = ndev;
  pl->patch_size = P;
  pl->serial = ++g_plan_serial;
  m.plan = pl;
  return true;
}

std::string ensure_device(Model& m, int device, std::shared_ptr<DeviceState>& out) {
  std::string err = ensure_plan(m);
  if (!err.empty()) return err;
  for (auto& d : m.dev)
    if (d->device == device) {
      if (d->plan_serial != m.plan->serial) {  // the model was re-planned: refresh the mirror (on its own device)
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != device) cudaSetDevice(device);
        err = d->ps.upload_from(m.plan->ps);
        if (cur != device) cudaSetDevice(cur);
        if (!err.empty()) return err;
        d->plan_serial = m.plan->serial;
      }
      out = d;
      return "";
    }
  if (cudaSetDevice(device) != cudaSuccess) return std::string("cudaSetDevice failed: ") + cudaGetErrorString(cudaGetLastError());
  auto ds = std::make_shared<DeviceState>();
  ds->device = device;
  err = ds->ps.upload_from(m.plan->ps);
  if (!err.empty()) return err;
  ds->plan_serial = m.plan->serial;
  m.dev.push_back(ds);
  out = ds;
  return "";
}

void drop_device_state(Model& m) {
  for (auto& d : m.dev) {
    cudaSetDevice(d->device);
    d.reset();
  }
  m.dev.clear();
  m.block_ctx.reset();
}

// rows of a shard: internal DOF range of its patches (+ the V-coil rows on the last shard)
void shard_rows(const Model& m, int nshards, int shard, int& p0, int& p1, std::vector<int>& row_ids, bool sym) {
  const PatchSet& ps = m.plan->ps;
  if (sym) shard_range_sym(ps, nshards, shard, p0, p1);
  else shard_range(ps, nshards, shard, p0, p1);
  row_ids.clear();
  for (int i = ps.patch_dof_ptr[p0]; i < ps.patch_dof_ptr[p1]; i++) row_ids.push_back(ps.dof_orig[i]);
  if (shard == nshards - 1)
    for (int j = 0; j < m.n_vcoils; j++) row_ids.push_back(m.np_active + m.nholes + j);
}

// Bands of a shard: contiguous patch ranges.  A band's rows are complete (all columns of this and later shards) as soon
// as its own tiles and the transposed copies from the earlier bands are done, so the rows of band k can leave the device
// while band k+1 is being built.  What stays exposed at the end is the copy of the LAST band, so it is made as thin as
// the tile count allows (a band of w row patches at the bottom of the matrix holds w^2/2 tiles and must still feed every
// SM: >= ~300 tiles); the patches before it are split into bands of (nearly) equal row counts.
std::vector<int> band_cuts(const PatchSet& ps, int p0, int p1, int nbands) {
  std::vector<int> cuts{p0};
  nbands = std::max(1, nbands);
  int plast = p1;  // first patch of the last band
  if (nbands > 1) {
    const int w = std::max((int)std::ceil(std::sqrt(2.0 * 300.0)), (p1 - p0) / 16);
    plast = p1 - w;
    if (plast <= p0 + 1) plast = p1, nbands = 1;  // too few patches for bands
  }
  const int i0 = ps.patch_dof_ptr[p0], i1 = ps.patch_dof_ptr[plast];
  const int nfirst = nbands > 1 ? nbands - 1 : 1;
  for (int b = 1; b < nfirst; b++) {
    const long long target = i0 + (long long)(i1 - i0) * b / nfirst;
    int p = cuts.back();
    while (p < plast && ps.patch_dof_ptr[p] < target) p++;
    if (p > cuts.back() && p < plast) cuts.push_back(p);
  }
  if (plast < p1 && plast > cuts.back()) cuts.push_back(plast);
  cuts.push_back(p1);
  return cuts;
}

// Number of bands of a single-device build whose rows go to the host.  Every band costs one kernel tail (about half a tile:
// 74 / ntiles of the build); the bands before the last one only have to keep the copy engine busy, so a handful is enough
// (the copy of the whole matrix takes 0.33 of the build on the 20k-vertex vessel, 0.85 at 100k vertices).
int auto_bands(const PatchSet& ps, int p0, int p1) {
  if (const char* e = std::getenv("THINCURR_B200_BANDS")) return std::max(1, std::atoi(e));
  const double np = p1 - p0;
  const double ntiles = np * (np + 1) / 2;
  // (a band cannot finish before its longest tile -- a near-field-heavy diagonal tile -- so small builds get few bands)
  if (ntiles < 2400) return 1;
  return ntiles < 6000 ? 2 : 6;
}

// Build the self-inductance rows of one shard on the current device into d_out[nrows][ld]
// (zeroed here).  Asynchronous on `stream` unless stats are requested.
// sym: symmetric partition, only the blocks against this and later shards are computed (upper trapezoid)
// nbands > 1: the shard is built band by band (same tiles, same bits); after band k `band_done(k, r0, r1)` is called with
// the band's row range [r0, r1) of d_out, whose rows are final once the work enqueued so far on `stream` has run.
std::string lmat_shard_device(Model& m, int nshards, int shard, double* d_out, long long ld, cudaStream_t stream,
                              unsigned long long* stats, bool sym, int nbands,
                              const std::function<std::string(int, int, int)>& band_done) {
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return "No CUDA device available (there is no CPU fallback)";
  std::shared_ptr<DeviceState> ds;
  std::string err = ensure_plan(m, nshards);
  if (!err.empty()) return err;
  err = ensure_device(m, device, ds);
  if (!err.empty()) return err;
  if (m.n_vcoils > 0 && !m.have_coil_mutuals) return "Coil mutuals required if, # of Vcoils > 0";
  const PatchSet& ps = m.plan->ps;
  int p0, p1;
  std::vector<int> row_ids;
  shard_rows(m, nshards, shard, p0, p1, row_ids, sym);
  const int I0 = ps.patch_dof_ptr[p0];
  const std::vector<int> cuts = band_cuts(ps, p0, p1, std::max(1, nbands));
  unsigned long long acc[8] = {0, 0, 0, 0, 0, 0, 0, ~0ull};
  for (size_t b = 0; b + 1 < cuts.size(); b++) {
    const int b0 = cuts[b], b1 = cuts[b + 1];
    const int i0 = ps.patch_dof_ptr[b0], i1 = ps.patch_dof_ptr[b1];
    double* d_band = d_out + (size_t)(i0 - I0) * ld;
    std::vector<int> row_out(ps.ndof, -1);
    for (int i = i0, r = 0; i < i1; i++, r++) row_out[i] = r;
    std::vector<Tile> tiles;
    build_self_tiles(ps, b0, b1, tiles, sym, b > 0 ? p0 : -1);
    if (cudaMemsetAsync(d_band, 0, (size_t)(i1 - i0) * ld * sizeof(double), stream) != cudaSuccess)
      return "cudaMemsetAsync failed on the output block";
    unsigned long long st[8] = {0};
    err = gpu_lmat_tiles(ds->ps, ds->ps, tiles, row_out, true, d_band, ld, stream, stats ? st : nullptr);
    if (!err.empty()) return err;
    if (stats) {
      for (int k = 0; k < 4; k++) acc[k] += st[k];
      acc[5] += st[4] - st[6];            // kernel time (ns) summed over the bands
      if (b == 0) acc[6] = st[6];         // start of the first band's kernel
      acc[4] = st[4];                     // end of the last band's kernel
      acc[7] = std::min(acc[7], st[7] - st[6]);
    }
    if (b > 0) {  // transposed blocks of the earlier bands of this shard (thin_wall.F90:1146-1151)
      err = gpu_symmetrize_cross(ds->ps, i0, i1, I0, i0, d_band, d_out, ld, stream);
      if (!err.empty()) return err;
    }
    if (band_done) {
      err = band_done((int)b, i0 - I0, i1 - I0);
      if (!err.empty()) return err;
    }
  }
  if (m.n_vcoils > 0) {
    const size_t nr0 = (size_t)(ps.patch_dof_ptr[p1] - I0);
    if (row_ids.size() > nr0 &&
        cudaMemsetAsync(d_out + nr0 * ld, 0, (row_ids.size() - nr0) * (size_t)ld * sizeof(double), stream) != cudaSuccess)
      return "cudaMemsetAsync failed on the output block";
    err = gpu_fill_vcoil_block(m, row_ids, d_out, ld, stream);
    if (!err.empty()) return err;
  }
  if (stats) {  // [6] -> [4]: summed duration of the band kernels; [7]: earliest first-CTA finish of a band
    acc[4] = acc[6] + acc[5];
    acc[7] = acc[6] + acc[7];
    std::memcpy(stats, acc, sizeof acc);
    stats[5] = 0;
  }
  return "";
}

}  // namespace tw

// =============================================================================================
// Block 1: reference-compatible entry points
// =============================================================================================
extern "C" {

static void (*g_abort_callback)(void) = nullptr;

void oftpy_init(int nthreads, bool quiet, const char* input_file, int* slens, void* abort_callback) {
  (void)nthreads; (void)input_file;
  g_abort_callback = (void (*)(void))abort_callback;  // installed by the Python layer (_interface.py:90-96)
  if (slens) {
    slens[0] = 4;   // OFT_MPI_PLEN (src/CMakeLists.txt:34-37)
    slens[1] = 80;  // OFT_SLEN
    slens[2] = THINCURR_PATH_SLEN;
    slens[3] = THINCURR_ERROR_SLEN;
  }
  if (!quiet) std::printf("thincurr-b200: CUDA operator-build backend (%d device(s) visible)\n", visible_devices());
}
void oftpy_set_nthreads(int nthreads) { (void)nthreads; }
void oftpy_set_debug(int debug_level) { g_debug = debug_level; }

void oftpy_load_xml(const char* xml_file, void** oft_node_ptr) {
  // silently leaves the pointer untouched on failure, like the reference (oft_base_f.F90:117-133)
  std::string err;
  auto root = xml_parse_file(cstr(xml_file), err);
  if (!root) return;
  auto* doc = new XmlDoc();
  doc->root = std::move(root);
  *oft_node_ptr = doc;
}

static const XmlNode* thincurr_node(void* xml_ptr, std::string& err) {
  if (!xml_ptr) return nullptr;
  auto* doc = (XmlDoc*)xml_ptr;
  const XmlNode* tc = doc->root->tag == "thincurr" ? doc->root.get() : doc->root->child("thincurr");
  if (!tc) err = "Error getting ThinCurr XML node";
  return tc;
}

static void fill_sizes(const Model& m, int* sizes) {
  int v[9] = {m.np, m.ne, m.nc, m.nreg, m.np_active, m.nholes, m.n_vcoils, m.nelems, m.n_icoils};
  std::memcpy(sizes, v, sizeof v);
}

void thincurr_setup(const char* mesh_file, int np, const double* r_loc, int nc, const int* lc_loc, const int* reg_loc,
                    const int* pmap_loc, int jumper_start, void** tw_ptr, int* sizes, char* error_str, void* xml_ptr) {
  set_err(error_str, "");
  std::string err;
  const XmlNode* tc = thincurr_node(xml_ptr, err);
  if (!err.empty()) return set_err(error_str, err);
  auto* m = new Model();
  std::vector<std::vector<int>> holes;
  std::vector<int> closures;
  if (np > 0) {
    // array mode: no node/side sets (thincurr_f.F90:77-98)
    err = m->setup_from_arrays(np, r_loc, nc, lc_loc, reg_loc, nullptr, holes, closures, tc);
  } else {
    NativeMesh nm;
    err = read_native_mesh(cstr(mesh_file), nm);
    if (err.empty()) {
      int nsets = (int)nm.nodesets.size(), nholes = nsets;
      if (jumper_start != 0) {  // thincurr_f.F90:173-190
        if (std::abs(jumper_start) > nsets) err = "\"jumper_start\" exceeds number of nodesets in file";
        int js = jumper_start < 0 ? nsets + 1 + jumper_start : jumper_start;
        nholes = js - 1;
        m->n_jumper_sets = nsets - nholes;
      }
      for (int h = 0; h < nholes && err.empty(); h++) {
        holes.emplace_back();
        for (int v : nm.nodesets[h]) holes.back().push_back(v - 1);
      }
      if (!nm.sidesets.empty())
        for (int c : nm.sidesets[0]) closures.push_back(c - 1);
      const int* pm = nullptr;
      if (!nm.pmap.empty()) pm = nm.pmap.data();
      else if (pmap_loc && pmap_loc[0] >= 0) pm = pmap_loc;
      if (err.empty())
        err = m->setup_from_arrays(nm.np, nm.r.data(), nm.nc, nm.lc.data(), nm.reg.data(), pm, holes, closures, tc);
    }
  }
  if (!err.empty()) {
    delete m;
    return set_err(error_str, err);
  }
  *tw_ptr = m;
  fill_sizes(*m, sizes);
}

// Device -> host copies.  Page-locked destinations take the copy engine directly (cudaMemcpyAsync, asynchronous).
// Pageable ones (a Fortran or numpy array of the caller, or a matrix too large to pin) would turn every
// cudaMemcpyAsync into a blocking staged copy on the calling thread and serialise the devices, so they are served by a
// helper thread per device that drains a job queue through a pinned double buffer.
static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

class HostCopier {
 public:
  explicit HostCopier(int device) : device_(device) {}
  ~HostCopier() { finish(); }
  // copy `bytes` from device memory `src` to pageable host memory `dst` once `after` (event, may be null) has completed
  void push(void* dst, const void* src, size_t bytes, cudaEvent_t after) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      jobs_.push_back(Job{dst, src, bytes, after});
      if (!started_) {
        started_ = true;
        th_ = std::thread([this] { run(); });
      }
    }
    cv_.notify_one();
  }
  std::string finish() {
    if (started_) {
      {
        std::lock_guard<std::mutex> lk(mu_);
        done_ = true;
      }
      cv_.notify_one();
      if (th_.joinable()) th_.join();
      started_ = false;
    }
    return err_;
  }

 private:
  struct Job {
    void* dst;
    const void* src;
    size_t bytes;
    cudaEvent_t after;
  };
  static constexpr size_t kSlot = (size_t)32 << 20;
  void run() {
    void* slot[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    bool ok = cudaSetDevice(device_) == cudaSuccess && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++)
      ok = cudaMallocHost(&slot[i], kSlot) == cudaSuccess && cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) err_ = "Pinned staging buffers for the device->host copy could not be set up";
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return done_ || !jobs_.empty(); });
        if (jobs_.empty()) break;
        j = jobs_.front();
        jobs_.pop_front();
      }
      if (!err_.empty()) continue;
      if (j.after && cudaStreamWaitEvent(st, j.after, 0) != cudaSuccess) err_ = "cudaStreamWaitEvent failed";
      size_t off = 0, po[2] = {0, 0}, pn[2] = {0, 0};
      int k = 0;
      auto flush = [&](int i) {
        if (!pn[i]) return;
        if (cudaEventSynchronize(ev[i]) != cudaSuccess) err_ = "Device->host copy failed";
        else std::memcpy((char*)j.dst + po[i], slot[i], pn[i]);
        pn[i] = 0;
      };
      while (off < j.bytes && err_.empty()) {
        const size_t n = std::min(kSlot, j.bytes - off);
        flush(k);
        if (cudaMemcpyAsync(slot[k], (const char*)j.src + off, n, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaEventRecord(ev[k], st) != cudaSuccess)
          err_ = std::string("Device->host copy failed: ") + cudaGetErrorString(cudaGetLastError());
        po[k] = off;
        pn[k] = n;
        off += n;
        k ^= 1;
      }
      flush(k);
      flush(k ^ 1);
    }
    for (int i = 0; i < 2; i++) {
      if (slot[i]) cudaFreeHost(slot[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
    if (st) cudaStreamDestroy(st);
  }
  int device_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Job> jobs_;
  std::thread th_;
  bool started_ = false, done_ = false;
  std::string err_;
};

struct DeviceGuard {  // restores the caller's current device
  int dev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&dev) != cudaSuccess) {
      cudaGetLastError();
      dev = -1;
    }
  }
  ~DeviceGuard() {
    if (dev >= 0) cudaSetDevice(dev);
  }
};

// devices a reference-facing operator call may use: all visible ones (THINCURR_B200_NDEV caps the count), or only the
// caller's current device when the process is one rank of a one-process-per-GPU job (torchrun / mpirun set LOCAL_RANK) or
// THINCURR_B200_ONE_DEVICE is set
static std::vector<int> build_devices() {
  const int n = visible_devices();
  std::vector<int> out;
  const bool one = std::getenv("THINCURR_B200_ONE_DEVICE") || (std::getenv("LOCAL_RANK") && !std::getenv("THINCURR_B200_NDEV"));
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess) {
    cudaGetLastError();
    cur = 0;
  }
  if (one || n <= 1) {
    if (n >= 1) out.push_back(cur);
    return out;
  }
  for (int g = 0; g < n; g++) out.push_back(g);
  return out;
}

// row-block scratch of the host-buffer entry points, kept between calls (a cudaMalloc/cudaFree pair of tens of GB per
// call costs more than the copies it serves); released by thincurr_b200_release_device / thincurr_b200_destroy
static std::string scratch_rows(DeviceState& ds, size_t bytes, double** out) {
  if (ds.scratch && ds.scratch_bytes >= bytes) {
    *out = ds.scratch;
    return "";
  }
  if (ds.scratch) cudaFree(ds.scratch);
  ds.scratch = nullptr;
  ds.scratch_bytes = 0;
  if (cudaMalloc((void**)&ds.scratch, std::max<size_t>(bytes, 8)) != cudaSuccess) {
    cudaGetLastError();
    return "Device allocation of the row block failed";
  }
  ds.scratch_bytes = bytes;
  *out = ds.scratch;
  return "";
}

// several devices stream a build only if every one of them can hold the whole matrix next to what it holds already
static bool stream_devices_ok(const std::vector<int>& devs_ids, size_t bytes) {
  if (std::getenv("THINCURR_B200_NO_MULTI_STREAM")) return false;
  DeviceGuard guard;
  for (int d : devs_ids) {
    size_t fr = 0, tot = 0;
    if (cudaSetDevice(d) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    if (tot < bytes + ((size_t)8 << 30)) return false;  // (the row-block scratch of an earlier call is reused or replaced)
  }
  return true;
}

// Streamed build into a page-locked host matrix (banded plan, see stream_ref_cuts).  A device holds the matrix in the
// reference layout and ONE launch of the tile kernel works through its bands (rows of a band against this and all later
// bands).  As soon as the tiles of a band are done the kernel's own CTAs run its mirror pass between two tiles and raise
// the band's flag in mapped host memory; this thread then hands the band's L-shaped part of the matrix -- its rows from
// its first column on, and its columns of all later rows -- to the copy engine as two strided copies, while the later
// bands are evaluated.  Every entry crosses the link exactly once and leaves as soon as the band that evaluated it (or
// its transposed twin) is done, so the link works in step with the evaluation from the first band on.  (A build whose
// rows may only leave complete -- all columns -- cannot start copying before its most expensive rows are finished and
// ends link-bound; separate launches per band end on their longest tile each.)
// Several devices: a band needs nothing from any other band, so the bands are dealt out to the devices by cost (no
// exchange, no peer access; every pair integral is still evaluated once) and every device streams its bands over its own
// link.
static std::string lmat_stream_host(Model& m, double* dst, const std::vector<int>& devs_ids, bool trace) {
  const auto tr0 = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  const Plan& pl = *m.plan;
  const PatchSet& ps = pl.ps;
  const size_t N = (size_t)m.nelems;
  const int nb = (int)pl.band_ref_ptr.size() - 1, ndev = (int)devs_ids.size();
  std::string err;
  auto ck = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && err.empty()) err = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaSuccess;
  };
  // band flags: a small mapped page-locked buffer kept for the life of the process (allocating and freeing page-locked
  // memory synchronises the devices); streamed builds of one process take turns
  constexpr int kMaxBands = 256;
  static int* g_flags = nullptr;
  static std::mutex g_stream_mu;
  std::lock_guard<std::mutex> stream_lock(g_stream_mu);
  if (!g_flags && !ck(cudaHostAlloc((void**)&g_flags, kMaxBands * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable), "cudaHostAlloc")) g_flags = nullptr;
  if (nb > kMaxBands) err = "Internal error: too many bands";
  if (!err.empty()) return err;
  std::memset(g_flags, 0, (size_t)nb * sizeof(int));

  // tiles of every band (most expensive first) and the bands of every device: greedy by cost, most expensive band first
  std::vector<std::vector<Tile>> band_tiles(nb);
  std::vector<double> band_cost(nb, 0.0);
  for (int b = 0; b < nb; b++) {
    build_self_tiles(ps, pl.band_patch_ptr[b], pl.band_patch_ptr[b + 1], band_tiles[b], false, b > 0 ? 0 : -1);
    for (const Tile& t : band_tiles[b]) band_cost[b] += t.cost;
  }
  std::vector<std::vector<int>> dev_bands(ndev);
  {
    std::vector<int> order(nb);
    for (int b = 0; b < nb; b++) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return band_cost[a] > band_cost[c]; });
    std::vector<double> load(ndev, 0.0);
    for (int b : order) {
      int g = 0;
      for (int k = 1; k < ndev; k++)
        if (load[k] < load[g]) g = k;
      dev_bands[g].push_back(b);
      load[g] += band_cost[b];
    }
    for (auto& v : dev_bands) std::sort(v.begin(), v.end());
  }
  std::vector<int> ref_patch(N, 0), row_out(ps.ndof, -1);
  for (int p = 0; p < ps.npatch; p++)
    for (int i = ps.patch_dof_ptr[p]; i < ps.patch_dof_ptr[p + 1]; i++) ref_patch[ps.dof_orig[i]] = p;
  for (int i = 0; i < ps.ndof; i++) row_out[i] = ps.dof_orig[i];  // output row = reference id
  if (trace) std::fprintf(stderr, "[lmat_stream_host] %d device(s), %d bands, tiles built at %.1f ms\n", ndev, nb, since(tr0));

  struct Dev {
    std::shared_ptr<DeviceState> ds;
    double* d = nullptr;
    cudaStream_t s = nullptr, cs = nullptr;
    int* d_ints = nullptr;
    int* flags = nullptr;  // this device's slice of the flag buffer (local band index)
    size_t next = 0;       // next band to hand to the copy engine
    bool done = false;
  };
  std::vector<Dev> devs(ndev);
  int flag_off = 0;
  // ---- every device: model upload, matrix, launch
  for (int g = 0; g < ndev && err.empty(); g++) {
    Dev& D = devs[g];
    const std::vector<int>& mine = dev_bands[g];
    const int nbl = (int)mine.size();
    D.flags = g_flags + flag_off;
    flag_off += nbl;
    if (nbl == 0) {
      D.done = true;
      continue;
    }
    if (!ck(cudaSetDevice(devs_ids[g]), "cudaSetDevice")) break;
    err = ensure_device(m, devs_ids[g], D.ds);
    if (!err.empty()) break;
    err = D.ds->ps.upload_from(ps);  // the call's inputs: host model -> device, every call
    if (!err.empty()) break;
    D.ds->plan_serial = pl.serial;
    err = scratch_rows(*D.ds, N * N * 8, &D.d);
    if (!err.empty()) break;
    if (!ck(cudaStreamCreateWithFlags(&D.s, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !ck(cudaStreamCreateWithFlags(&D.cs, cudaStreamNonBlocking), "cudaStreamCreate"))
      break;
    // Queue order: latest start time first.  A band should be complete when the work of this device's bands up to it is
    // done (deadline = that work spread over the SMs, in units of the tile cost model); a tile must start its own cost
    // before the deadline of its band, so the long tiles of a band (its near-field diagonal blocks) start ahead of the
    // cheap tiles of the band before, and the queue ends on cheap tiles.
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, devs_ids[g]);
    std::vector<Tile> tiles;
    std::vector<int> tile_band;
    std::vector<double> key;
    double deadline = 0.0;
    for (int l = 0; l < nbl; l++) {
      const int b = mine[l];
      deadline += band_cost[b] / nsm;
      for (const Tile& t : band_tiles[b]) {
        tiles.push_back(t);
        tile_band.push_back(l);
        key.push_back(deadline - t.cost);
      }
    }
    {
      std::vector<int> order(tiles.size());
      for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return key[a] < key[c]; });
      std::vector<Tile> t2(tiles.size());
      std::vector<int> b2(tiles.size());
      for (size_t i = 0; i < order.size(); i++) {
        t2[i] = tiles[order[i]];
        b2[i] = tile_band[order[i]];
      }
      tiles.swap(t2);
      tile_band.swap(b2);
    }
    // band arrays of the kernel: tiles per band [nbl], first / one-past-last reference id of the rows [2 nbl], then 3 nbl counters
    std::vector<int> hb(3 * nbl);
    for (int l = 0; l < nbl; l++) {
      hb[l] = (int)band_tiles[mine[l]].size();
      hb[nbl + 2 * l] = pl.band_ref_ptr[mine[l]];
      hb[nbl + 2 * l + 1] = pl.band_ref_ptr[mine[l] + 1];
    }
    const size_t nt = tiles.size();
    ck(cudaMallocAsync((void**)&D.d_ints, (N + nt + 6 * (size_t)nbl) * sizeof(int), D.s), "cudaMallocAsync");
    ck(cudaMemcpyAsync(D.d_ints, ref_patch.data(), N * sizeof(int), cudaMemcpyHostToDevice, D.s), "cudaMemcpyAsync");
    ck(cudaMemcpyAsync(D.d_ints + N, tile_band.data(), nt * sizeof(int), cudaMemcpyHostToDevice, D.s), "cudaMemcpyAsync");
    ck(cudaMemcpyAsync(D.d_ints + N + nt, hb.data(), hb.size() * sizeof(int), cudaMemcpyHostToDevice, D.s), "cudaMemcpyAsync");
    ck(cudaMemsetAsync(D.d_ints + N + nt + 3 * nbl, 0, (size_t)3 * nbl * sizeof(int), D.s), "cudaMemsetAsync");
    ck(cudaStreamSynchronize(D.s), "cudaStreamSynchronize");  // (host vectors are locals; nothing large is queued yet)
    ck(cudaMemsetAsync(D.d, 0, N * N * 8, D.s), "cudaMemsetAsync");
    if (!err.empty()) break;
    StreamBands sb;
    sb.nbands = nbl;
    sb.N = (int)N;
    sb.flags = D.flags;
    sb.d_ref_patch = D.d_ints;
    sb.d_tile_band = D.d_ints + N;
    sb.d_bands = D.d_ints + N + nt;
    err = gpu_lmat_tiles(D.ds->ps, D.ds->ps, tiles, row_out, true, D.d, (long long)N, D.s, nullptr, nullptr, false, &sb);
    if (trace) std::fprintf(stderr, "[lmat_stream_host] device %d: %d bands, %zu tiles, launched at %.1f ms\n", devs_ids[g], nbl, tiles.size(), since(tr0));
  }
  // ---- hand the finished bands to the copy engines
  for (bool pending = err.empty(); pending && err.empty();) {
    pending = false;
    bool progress = false;
    for (int g = 0; g < ndev && err.empty(); g++) {
      Dev& D = devs[g];
      if (D.done) continue;
      pending = true;
      volatile int* vf = D.flags;
      if (!vf[D.next]) {
        if (!ck(cudaSetDevice(devs_ids[g]), "cudaSetDevice")) break;
        const cudaError_t q = cudaStreamQuery(D.s);
        if (q == cudaErrorNotReady) continue;
        if (q != cudaSuccess) ck(q, "Kernel execution failed");
        else if (!vf[D.next]) err = "Internal error: the streamed build ended without finishing its bands";
        if (!err.empty()) break;
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      const int b = dev_bands[g][D.next];
      const int R0 = pl.band_ref_ptr[b], R1 = pl.band_ref_ptr[b + 1];
      const size_t o0 = (size_t)R0 * N + R0, o1 = (size_t)R1 * N + R0;
      if (!ck(cudaSetDevice(devs_ids[g]), "cudaSetDevice")) break;
      ck(cudaMemcpy2DAsync(dst + o0, N * 8, D.d + o0, N * 8, (N - R0) * 8, (size_t)(R1 - R0), cudaMemcpyDeviceToHost, D.cs), "Device->host copy");
      if ((size_t)R1 < N)
        ck(cudaMemcpy2DAsync(dst + o1, N * 8, D.d + o1, N * 8, (size_t)(R1 - R0) * 8, N - R1, cudaMemcpyDeviceToHost, D.cs), "Device->host copy");
      if (trace) std::fprintf(stderr, "[lmat_stream_host] device %d: band %d (rows [%d,%d), %zu tiles) final at %.1f ms\n", devs_ids[g], b, R0, R1, band_tiles[b].size(), since(tr0));
      progress = true;
      if (++D.next == dev_bands[g].size()) D.done = true;
    }
    if (pending && !progress && err.empty()) std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
  for (int g = 0; g < ndev; g++) {
    Dev& D = devs[g];
    if (!D.s) continue;
    cudaSetDevice(devs_ids[g]);
    if (D.d_ints) cudaFreeAsync(D.d_ints, D.s);
    cudaError_t ce = cudaStreamSynchronize(D.s);
    if (trace) std::fprintf(stderr, "[lmat_stream_host] device %d: build stream done at %.1f ms\n", devs_ids[g], since(tr0));
    if (ce == cudaSuccess && D.cs) ce = cudaStreamSynchronize(D.cs);
    if (trace) std::fprintf(stderr, "[lmat_stream_host] device %d: copies done at %.1f ms\n", devs_ids[g], since(tr0));
    if (ce != cudaSuccess && err.empty()) err = std::string("Kernel execution failed: ") + cudaGetErrorString(ce);
    cudaStreamDestroy(D.s);
    if (D.cs) cudaStreamDestroy(D.cs);
  }
  cudaGetLastError();
  if (trace) std::fprintf(stderr, "[lmat_stream_host] returning at %.1f ms\n", since(tr0));
  return err;
}

// full self-inductance matrix into host memory dst[nelems][nelems] (reference layout).  Large models without V-coils and
// with a page-locked destination: the streamed build above (lmat_stream_host).  Otherwise the rows are sharded over the
// usable devices, each shard copied straight to its place.  One device: the matrix is built in row bands (separate
// launches) and the rows of a finished band go to the host (copy engine, or the staging thread of a pageable
// destination) while the next band is evaluated.  Several devices that can read each other's memory: symmetric shards
// (no pair integral evaluated twice); all builds are launched first, then every device fetches the transposed blocks of
// the other shards over NVLink (symmetrize_cross_kernel on peer memory) and sends its rows to the host.
static std::string lmat_full_host(Model& m, double* dst) {
  const size_t N = (size_t)m.nelems;
  const bool trace = std::getenv("THINCURR_B200_TRACE") != nullptr;
  const auto tr0 = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  DeviceGuard guard;
  std::vector<int> devs_ids = build_devices();
  if (devs_ids.empty()) return "No CUDA device available (the B200 backend has no CPU fallback)";
  std::string err;
  const bool pinned_dst = is_pinned(dst);
  bool banded = false;
  if (m.n_vcoils == 0 && !std::getenv("THINCURR_B200_FULL_ROWS") && (devs_ids.size() == 1 || (pinned_dst && stream_devices_ok(devs_ids, N * N * 8))))
    banded = ensure_banded_plan(m, err, (int)devs_ids.size());
  if (banded) {
    // large model: banded plan; a page-locked destination is served by the streamed build (bands dealt out to the devices),
    // a pageable one (one device) by the row bands below on the same patches (same bits)
    if (pinned_dst) return lmat_stream_host(m, dst, devs_ids, trace);
  } else {
    if (!err.empty()) return err;
    // one device: the rows leave band by band and every band needs enough tiles for all SMs: patches as for 8 shards
    err = replan(m, devs_ids.size() == 1 ? 8 : (int)devs_ids.size());
    if (!err.empty()) return err;
  }
  const PatchSet& ps = m.plan->ps;
  int ndev = std::min((int)devs_ids.size(), std::max(1, ps.npatch));
  devs_ids.resize(ndev);
  bool sym = ndev > 1 && m.n_vcoils == 0 && !std::getenv("THINCURR_B200_FULL_ROWS");
  for (int g = 1; g < ndev && sym; g++)
    for (int s = 0; s < g && sym; s++) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devs_ids[g], devs_ids[s]) != cudaSuccess || !can) sym = false;
    }
  cudaGetLastError();
  struct Dev {
    double* d = nullptr;
    cudaStream_t s = nullptr, cs = nullptr;  // build stream, copy stream
    cudaEvent_t built = nullptr;
    std::vector<cudaEvent_t> band_ev;
    std::vector<int> rows;
    int i0 = 0, i1 = 0;  // internal DOF range of the rows
    std::shared_ptr<DeviceState> ds;
  };
  std::vector<Dev> devs(ndev);
  const bool pinned = is_pinned(dst);
  size_t ncopies = 0;
  std::vector<std::unique_ptr<HostCopier>> copier(ndev);
  // rows [r0,r1) of device g to their place in the reference layout Lmat(:,row), runs of consecutive reference ids as
  // one copy; the copies wait for `after` (an event of the build stream)
  auto rows_to_host = [&](int g, int r0, int r1, cudaEvent_t after) -> std::string {
    Dev& D = devs[g];
    if (pinned && cudaStreamWaitEvent(D.cs, after, 0) != cudaSuccess) return "cudaStreamWaitEvent failed";
    if (!pinned && !copier[g]) copier[g].reset(new HostCopier(devs_ids[g]));
    for (int r = r0; r < r1;) {
      int e = r + 1;
      while (e < r1 && D.rows[e] == D.rows[e - 1] + 1) e++;
      double* to = dst + (size_t)D.rows[r] * N;
      const double* from = D.d + (size_t)r * N;
      const size_t bytes = (size_t)(e - r) * N * 8;
      if (!pinned) copier[g]->push(to, from, bytes, after);
      else if (cudaMemcpyAsync(to, from, bytes, cudaMemcpyDeviceToHost, D.cs) != cudaSuccess)
        return std::string("Device->host copy failed: ") + cudaGetErrorString(cudaGetLastError());
      ncopies++;
      r = e;
    }
    return "";
  };
  auto ck = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && err.empty()) err = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaSuccess;
  };
  const double tr_plan = since(tr0);
  // ---- phase 1: every device starts building its shard
  for (int g = 0; g < ndev && err.empty(); g++) {
    Dev& D = devs[g];
    if (!ck(cudaSetDevice(devs_ids[g]), "cudaSetDevice")) break;
    int p0, p1;
    shard_rows(m, ndev, g, p0, p1, D.rows, sym);
    D.i0 = ps.patch_dof_ptr[p0];
    D.i1 = ps.patch_dof_ptr[p1];
    if (D.rows.empty()) continue;
    if (!ck(cudaStreamCreateWithFlags(&D.s, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !ck(cudaStreamCreateWithFlags(&D.cs, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !ck(cudaEventCreateWithFlags(&D.built, cudaEventDisableTiming), "cudaEventCreate"))
      break;
    err = ensure_device(m, devs_ids[g], D.ds);
    if (!err.empty()) break;
    err = D.ds->ps.upload_from(ps);  // the call's inputs: host model -> device, every call
    if (!err.empty()) break;
    D.ds->plan_serial = m.plan->serial;
    err = scratch_rows(*D.ds, D.rows.size() * N * 8, &D.d);  // row block kept between calls
    if (!err.empty()) break;
    const int nb = ndev == 1 ? auto_bands(ps, p0, p1) : 1;
    auto band_done = [&, g](int, int r0, int r1) -> std::string {
      // (single device) the band's rows are final: hand them to the copy engine behind an event
      cudaEvent_t ev = nullptr;
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return "cudaEventCreate failed";
      devs[g].band_ev.push_back(ev);
      if (cudaEventRecord(ev, devs[g].s) != cudaSuccess) return "cudaEventRecord failed";
      return rows_to_host(g, r0, r1, ev);
    };
    if (ndev == 1 && m.n_vcoils == 0) {
      err = lmat_shard_device(m, ndev, g, D.d, (long long)N, D.s, nullptr, sym, nb, band_done);
    } else {
      err = lmat_shard_device(m, ndev, g, D.d, (long long)N, D.s, nullptr, sym);
      if (err.empty()) ck(cudaEventRecord(D.built, D.s), "cudaEventRecord");
    }
  }
  // ---- phase 2: transposed blocks of the earlier shards from peer memory, then the rows go to the host
  if (!(ndev == 1 && m.n_vcoils == 0)) {
    for (int g = 0; g < ndev && err.empty(); g++) {
      Dev& D = devs[g];
      if (!D.d) continue;
      if (!ck(cudaSetDevice(devs_ids[g]), "cudaSetDevice")) break;
      for (int k = 1; k < ndev && sym && err.empty(); k++) {  // rotated order: every device reads from a different peer
        const int s = (g + k) % ndev;
        if (!devs[s].d) continue;
        cudaError_t pe = cudaDeviceEnablePeerAccess(devs_ids[s], 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
          ck(pe, "cudaDeviceEnablePeerAccess");
          break;
        }
        cudaGetLastError();
        if (!ck(cudaStreamWaitEvent(D.s, devs[s].built, 0), "cudaStreamWaitEvent")) break;
        err = gpu_symmetrize_cross(D.ds->ps, D.i0, D.i1, devs[s].i0, devs[s].i1, D.d, devs[s].d, (long long)N, D.s, true);
      }
    }
    for (int g = 0; g < ndev && err.empty(); g++) {
      Dev& D = devs[g];
      if (!D.d) continue;
      if (!ck(cudaSetDevice(devs_ids[g]), "cudaSetDevice")) break;
      cudaEvent_t ev = nullptr;
      if (!ck(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate")) break;
      D.band_ev.push_back(ev);
      if (!ck(cudaEventRecord(ev, D.s), "cudaEventRecord")) break;
      err = rows_to_host(g, 0, (int)D.rows.size(), ev);
    }
  }
  for (int g = 0; g < ndev; g++) {
    if (!devs[g].s) continue;
    cudaSetDevice(devs_ids[g]);
    cudaError_t ce = cudaStreamSynchronize(devs[g].s);
    if (trace) std::fprintf(stderr, "[lmat_full_host] device %d: build stream done at %.1f ms\n", devs_ids[g], since(tr0));
    if (ce == cudaSuccess && devs[g].cs) ce = cudaStreamSynchronize(devs[g].cs);
    if (trace) std::fprintf(stderr, "[lmat_full_host] device %d: copies done at %.1f ms\n", devs_ids[g], since(tr0));
    if (ce != cudaSuccess && err.empty()) err = std::string("Kernel execution failed: ") + cudaGetErrorString(ce);
    if (copier[g]) {
      std::string ce2 = copier[g]->finish();
      if (err.empty()) err = ce2;
    }
  }
  if (trace)
    std::fprintf(stderr, "[lmat_full_host] %d device(s), %s destination, plan %.1f ms, total %.1f ms, %zu row-run copies, sym %d\n", ndev,
                 pinned ? "pinned" : "pageable", tr_plan, since(tr0), ncopies, (int)sym);
  for (int g = 0; g < ndev; g++) {  // (all devices are done reading each other's blocks)
    cudaSetDevice(devs_ids[g]);
    if (devs[g].s) cudaStreamDestroy(devs[g].s);
    if (devs[g].cs) cudaStreamDestroy(devs[g].cs);
    if (devs[g].built) cudaEventDestroy(devs[g].built);
    for (cudaEvent_t e : devs[g].band_ev) cudaEventDestroy(e);
  }
  cudaGetLastError();
  return err;
}

void thincurr_Lmat(void* tw_ptr, bool use_hodlr, void** Lmat_ptr, const char* cache_file, char* error_str) {
  Model& m = *(Model*)tw_ptr;
  if (m.n_vcoils > 0 && !m.have_coil_mutuals) return set_err(error_str, "Coil mutuals required if, # of Vcoils > 0");
  set_err(error_str, "");
  if (use_hodlr) return set_err(error_str, "HODLR compression is not provided by the B200 dense backend");
  std::string cache = cstr(cache_file);
  const size_t N = (size_t)m.nelems;
  if (!cache.empty() && cache != "none" && lmat_cache_read(m, cache)) {
    *Lmat_ptr = m.Lmat.p;
    return;
  }
  std::printf(" Building element<->element self inductance matrix\n");
  auto t0 = std::chrono::steady_clock::now();
  if (visible_devices() < 1) return set_err(error_str, "No CUDA device available (the B200 backend has no CPU fallback)");
  m.Lmat.alloc(N * N, false);  // every entry is overwritten by the device->host copies
  if (!m.Lmat.p) return set_err(error_str, "Host allocation of the inductance matrix failed");
  std::string err = lmat_full_host(m, m.Lmat.p);
  if (!err.empty()) return set_err(error_str, err);
  double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("   Time = %s\n", time_to_string(el).c_str());
  if (!cache.empty() && cache != "none") lmat_cache_write(m, cache);
  *Lmat_ptr = m.Lmat.p;
}

void thincurr_cross_coupling(void* tw_ptr1, void* tw_ptr2, double* Mmat, const char* cache_file, char* error_str) {
  set_err(error_str, "");
  Model &m1 = *(Model*)tw_ptr1, &m2 = *(Model*)tw_ptr2;
  std::string cache = cstr(cache_file);
  if (!cache.empty() && cache != "none" && mutual_cache_read(m1, m2, Mmat, cache)) return;
  std::printf(" Building element<->element mutual inductance matrix\n");
  auto t0 = std::chrono::steady_clock::now();
  std::string err = gpu_cross_coupling(m1, m2, Mmat);
  if (!err.empty()) return set_err(error_str, err);
  double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("   Time = %s\n", time_to_string(el).c_str());
  if (!cache.empty() && cache != "none") mutual_cache_write(m1, m2, Mmat, cache);
}

void thincurr_Mcoil(void* tw_ptr, void** Mc_ptr, const char* cache_file, char* error_str) {
  set_err(error_str, "");
  Model& m = *(Model*)tw_ptr;
  std::string cache = cstr(cache_file);
  if (!(!cache.empty() && cache != "none" && mcoil_cache_read(m, cache))) {
    std::printf(" Building coil<->element inductance matrices\n");
    auto t0 = std::chrono::steady_clock::now();
    std::string err = gpu_mcoil(m);
    if (!err.empty()) return set_err(error_str, err);
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("   Time = %s\n", time_to_string(el).c_str());
    for (int i = 0; i < m.n_vcoils; i++)
      std::printf(" Vcoil %4d: L [H] = %12.4E\n", i + 1, m.vcoils[i].Lself * 1.e-7);
    if (!cache.empty() && cache != "none") mcoil_cache_write(m, cache);
  }
  *Mc_ptr = m.Ael2dr.p;
}

void thincurr_Msensor(void* tw_ptr, const char* sensor_file, void** Ms_ptr, void** Msc_ptr, int* nsensors, int* njumpers,
                      void** sensor_ptr, const char* cache_file, char* error_str) {
  set_err(error_str, "");
  Model& m = *(Model*)tw_ptr;
  if (*sensor_ptr) {
    delete (Sensors*)*sensor_ptr;
    *sensor_ptr = nullptr;
  }
  auto* sens = new Sensors();
  std::string err = read_floops(cstr(sensor_file), *sens);
  if (!err.empty()) {
    delete sens;
    return set_err(error_str, err);
  }
  std::string cache = cstr(cache_file);
  if (!(!cache.empty() && cache != "none" && msensor_cache_read(m, (int)sens->floops.size(), cache))) {
    std::printf(" Building element->sensor inductance matrix\n");
    err = gpu_msensor(m, *sens);
    if (!err.empty()) {
      delete sens;
      return set_err(error_str, err);
    }
    if (!cache.empty() && cache != "none") msensor_cache_write(m, (int)sens->floops.size(), cache);
  }
  *Ms_ptr = m.Ael2sen.p;
  *Msc_ptr = m.Adr2sen.p;
  *sensor_ptr = sens;
  *nsensors = (int)sens->floops.size();
  *njumpers = sens->njumpers;
}

void thincurr_get_sensor_name(void* sensor_ptr, int sensor_ind, char* sensor_name, char* error_str) {
  set_err(error_str, "");
  auto* s = (Sensors*)sensor_ptr;
  if (!s || sensor_ind < 1 || sensor_ind > (int)s->floops.size()) return set_err(error_str, "Invalid sensor index");
  std::snprintf(sensor_name, 40, "%s", s->floops[sensor_ind - 1].name.c_str());
}

void thincurr_Bmat(void* tw_ptr, void* hodlr_ptr, void** Bmat_ptr, void** Bdr_ptr, const char* cache_file, char* error_str) {
  set_err(error_str, "");
  if (hodlr_ptr) return set_err(error_str, "HODLR compression is not provided by the B200 dense backend");
  Model& m = *(Model*)tw_ptr;
  std::string cache = cstr(cache_file);
  const bool use_cache = !cache.empty() && cache != "none";
  if (!(use_cache && bmat_cache_read(m, cache))) {
    std::printf(" Building element->element magnetic reconstruction operator\n");
    std::string err = gpu_bmat(m);
    if (!err.empty()) return set_err(error_str, err);
    if (use_cache) bmat_cache_write(m, cache);
  }
  *Bmat_ptr = m.Bel.p;
  *Bdr_ptr = m.Bdr.p;
}

void thincurr_Rmat(void* tw_ptr, int** kr_ptr, int** lc_ptr, double** mat_ptr, char* error_str) {
  set_err(error_str, "");
  Model& m = *(Model*)tw_ptr;
  std::printf(" Building resistivity matrix\n");
  m.build_rmat();
  *kr_ptr = m.R_kr.data();
  *lc_ptr = m.R_lc.data();
  *mat_ptr = m.R_val.data();
}

void thincurr_get_eta(void* tw_ptr, double* eta_surf, char* error_str) {
  set_err(error_str, "");
  Model& m = *(Model*)tw_ptr;
  for (int i = 0; i < m.nreg; i++) eta_surf[i] = m.eta_surf[i] * kMu0;
}

void thincurr_set_eta(void* tw_ptr, const double* eta_surf, const double* eta_vol, const double* thickness, char* error_str) {
  // thincurr_f.F90:738-883: eta_surf alone, or any two (third derived), or all three (eta_surf recomputed)
  set_err(error_str, "");
  Model& m = *(Model*)tw_ptr;
  int cnt = (eta_surf != nullptr) + (eta_vol != nullptr) + (thickness != nullptr);
  if (cnt == 0)
    return set_err(error_str, "Provide eta_surf alone, or any two of \"eta_surf\", \"eta_vol\", and \"thickness\" to thincurr_set_eta");
  if (cnt == 1 && !eta_surf)
    return set_err(error_str, "\"eta_surf\" must be provided alone, or with one of \"eta_vol\" or \"thickness\"");
  for (int i = 0; i < m.nreg; i++) {
    if (eta_surf) m.eta_surf[i] = eta_surf[i] / kMu0;
    if (eta_vol) m.eta_vol[i] = eta_vol[i] / kMu0;
    if (thickness) m.thickness[i] = thickness[i];
    if (eta_vol && thickness) m.eta_surf[i] = m.eta_vol[i] / m.thickness[i];
    else if (eta_surf && thickness) m.eta_vol[i] = m.eta_surf[i] * m.thickness[i];
    else if (eta_surf && eta_vol) m.thickness[i] = m.eta_vol[i] / m.eta_surf[i];
  }
}

// ---- remaining names the reference's Python layer binds at import (ThinCurr/_interface.py:21-119) ---------------------
// Cheap model queries are answered natively; the dense apply and the iterative eigen solve run on the device
// (tw_solve.cu); everything that belongs to the reference's downstream solvers / plotting returns an error through
// error_str (or, for the wrappers without one, prints it and calls the abort callback oftpy_init received, the
// reference's oft_abort convention).
static const char* kNotProvided = "is not provided by the B200 operator-build backend (use the reference liboftpy for this step)";
static void no_errstr_abort(const char* name) {
  std::fprintf(stderr, "ERROR: %s %s\n", name, kNotProvided);
  if (g_abort_callback) g_abort_callback();
}
// names the reference's base package binds at import (OpenFUSIONToolkit/_interface.py:114-132); mesh objects of the
// other physics modules are not part of this backend
void oft_setup_smesh(int, int, const double*, int, int, const int*, const int*, int*, void** mesh_ptr) {
  if (mesh_ptr) *mesh_ptr = nullptr;
  no_errstr_abort("oft_setup_smesh");
}
void oft_smesh_get(void*, int*, int*, double**, int*, int*, int**, int**, int*, char* error_str) {
  set_err(error_str, std::string("oft_smesh_get ") + kNotProvided);
}
void oft_setup_vmesh(int, const double*, int, int, const int*, const int*, int*, void** mesh_ptr) {
  if (mesh_ptr) *mesh_ptr = nullptr;
  no_errstr_abort("oft_setup_vmesh");
}
void oft_vmesh_get(void*, int*, double**, int*, int*, int**, int**, int*, char* error_str) {
  set_err(error_str, std::string("oft_vmesh_get ") + kNotProvided);
}
void dump_cov(void) {}
void thincurr_scale_va(void* tw_ptr, double* vals, bool div_flag) {  // thincurr_f.F90:453-466
  Model& m = *(Model*)tw_ptr;
  for (int i = 0; i < m.np; i++) vals[i] = div_flag ? vals[i] / m.va[i] : vals[i] * m.va[i];
}
void thincurr_get_eta_vol(void* tw_ptr, double* eta_vol, char* error_str) {  // thincurr_f.F90:721-731
  set_err(error_str, "");
  Model& m = *(Model*)tw_ptr;
  for (int i = 0; i < m.nreg; i++) eta_vol[i] = m.eta_vol[i] * kMu0;
}
void thincurr_get_thickness(void* tw_ptr, double* thickness, char* error_str) {  // thincurr_f.F90:887-902
  set_err(error_str, "");
  Model& m = *(Model*)tw_ptr;
  for (int i = 0; i < m.nreg; i++) thickness[i] = m.thickness[i];
}
void thincurr_setup_io(void*, const char*, bool, bool, char* error_str) { set_err(error_str, std::string("thincurr_setup_io ") + kNotProvided); }
void thincurr_recon_curr(void*, const double*, double*, int) { no_errstr_abort("thincurr_recon_curr"); }
void thincurr_recon_field(void*, const double*, const double*, double*, void*) { no_errstr_abort("thincurr_recon_field"); }
void thincurr_save_field(void*, const double*, const char*) { no_errstr_abort("thincurr_save_field"); }
void thincurr_save_scalar(void*, const double*, const char*) { no_errstr_abort("thincurr_save_scalar"); }
void thincurr_curr_regmat(void*, double*, char* error_str) { set_err(error_str, std::string("thincurr_curr_regmat ") + kNotProvided); }
void thincurr_freq_response(void*, bool, int, double, double*, void*, char* error_str) {
  set_err(error_str, std::string("thincurr_freq_response ") + kNotProvided);
}
void thincurr_time_domain(void*, bool, double, int, double, double, bool, int, int, const double*, void*, int, const double*, int, const double*,
                          bool, void*, void*, char* error_str) {
  set_err(error_str, std::string("thincurr_time_domain ") + kNotProvided);
}
void thincurr_time_domain_plot(void*, bool, bool, int, int, void*, const double*, int, void*, char* error_str) {
  set_err(error_str, std::string("thincurr_time_domain_plot ") + kNotProvided);
}
void thincurr_reduce_model(void* tw_ptr, const char* filename, int neigs, const double* eig_vec, bool compute_B, void* sensor_ptr,
                           void* hodlr_ptr, char* error_str) {
  // thincurr_f.F90:1208-1247 -> tw_reduce_model (thin_wall_solvers.F90:1180-1359), dense L only
  Model& m = *(Model*)tw_ptr;
  if (m.nelems <= 0) return set_err(error_str, "Invalid ThinCurr model, may not be setup yet");
  if (hodlr_ptr) return set_err(error_str, "HODLR compression is not provided by the B200 dense backend");
  if (!m.Lmat.p) return set_err(error_str, "Inductance matrix required, but not computed");
  if (m.R_kr.empty()) return set_err(error_str, "Resistance matrix required, but not computed");
  set_err(error_str, "");
  std::string err = reduce_model(m, (const Sensors*)sensor_ptr, filename ? filename : "", neigs, eig_vec, compute_B);
  if (!err.empty()) set_err(error_str, err);
}
void thincurr_cross_eval(void* tw_ptr1, void* tw_ptr2, int nrhs, const double* vec1, double* vec2, char* error_str) {
  // thincurr_f.F90:525-541 -> tw_compute_Lmat_MF (thin_wall.F90:1190-1414)
  set_err(error_str, "");
  if (!tw_ptr1 || !tw_ptr2) return set_err(error_str, "Invalid ThinCurr model, may not be setup yet");
  Model &m1 = *(Model*)tw_ptr1, &m2 = *(Model*)tw_ptr2;
  if (m1.verbose) std::printf(" Applying MF element<->element inductance matrix\n");
  std::string err = gpu_cross_eval(m1, m2, nrhs, vec1, vec2, nullptr);
  if (!err.empty()) set_err(error_str, err);
}
void thincurr_apply_Lmat(void* tw_ptr, double* vals, void* hodlr_ptr) {
  Model& m = *(Model*)tw_ptr;
  if (hodlr_ptr || !m.Lmat.p) return no_errstr_abort("thincurr_apply_Lmat without a dense inductance matrix");
  std::string err = gpu_apply_host_matrix(m.Lmat.p, (size_t)m.nelems, (size_t)m.nelems, vals);
  if (!err.empty()) {
    std::fprintf(stderr, "ERROR: thincurr_apply_Lmat: %s\n", err.c_str());
    if (g_abort_callback) g_abort_callback();
  }
}
void thincurr_eigenvalues(void* tw_ptr, bool direct, int neigs, double* eig_vals, double* eig_vec, void* hodlr_ptr, char* error_str) {
  // thincurr_f.F90:975-1013.  Iterative path only (lr_eigenmodes_arpack, thin_wall_solvers.F90:119-224): Lanczos on the
  // device-resident L with R factorised on the host; the dense LAPACK path (direct=true) stays with the reference.
  Model& m = *(Model*)tw_ptr;
  if (m.nelems <= 0) return set_err(error_str, "Invalid ThinCurr model, may not be setup yet");
  if (hodlr_ptr) return set_err(error_str, "HODLR compression is not provided by the B200 dense backend");
  if (!m.Lmat.p) return set_err(error_str, "Inductance matrix required, but not computed");
  if (m.R_kr.empty()) return set_err(error_str, "Resistance matrix required, but not computed");
  if (direct) return set_err(error_str, std::string("thincurr_eigenvalues(direct=True) ") + kNotProvided);
  set_err(error_str, "");
  std::string err = gpu_lr_eigenmodes_host(m, neigs, eig_vals, eig_vec);
  if (!err.empty()) set_err(error_str, err);
}

// =============================================================================================
// Block 2: flat / sharded interface
// =============================================================================================
const char* thincurr_b200_last_error(void) { return g_last_error.c_str(); }
int thincurr_b200_device_count(void) { return visible_devices(); }

void thincurr_b200_destroy(void* tw_ptr) {
  if (!tw_ptr) return;
  auto* m = (Model*)tw_ptr;
  drop_device_state(*m);
  delete m;
}

int thincurr_b200_setup(int np, const double* r, int nc, const int* lc, const int* reg, const int* pmap, int nnodesets,
                        const int* nodeset_ptr, const int* nodeset_val, int nclosures, const int* closures, void* xml_ptr,
                        void** tw_ptr, int* sizes) {
  std::string err;
  const XmlNode* tc = thincurr_node(xml_ptr, err);
  if (!err.empty()) return fail(err);
  std::vector<std::vector<int>> holes(nnodesets);
  for (int h = 0; h < nnodesets; h++)
    for (int k = nodeset_ptr[h]; k < nodeset_ptr[h + 1]; k++) holes[h].push_back(nodeset_val[k] - 1);
  std::vector<int> cl;
  for (int k = 0; k < nclosures; k++) cl.push_back(closures[k] - 1);
  auto* m = new Model();
  err = m->setup_from_arrays(np, r, nc, lc, reg, pmap, holes, cl, tc);
  if (!err.empty()) {
    delete m;
    return fail(err);
  }
  *tw_ptr = m;
  if (sizes) fill_sizes(*m, sizes);
  return 0;
}

int thincurr_b200_model_from_tw(int np, const double* r, int nc, const int* lc, const int* reg, const int* pmap, int np_active,
                                int nholes, const int* kfh, const int* lfh, const double* ca, const double* qbasis,
                                void** tw_ptr) {
  auto* m = new Model();
  std::string err = m->setup_from_tw(np, r, nc, lc, reg, pmap, np_active, nholes, kfh, lfh, ca, qbasis);
  if (!err.empty()) {
    delete m;
    return fail(err);
  }
  *tw_ptr = m;
  return 0;
}

int thincurr_b200_Lmat_host(void* tw_ptr, double* Lmat) {
  Model& m = *(Model*)tw_ptr;
  if (m.n_vcoils > 0 && !m.have_coil_mutuals) return fail("Coil mutuals required if, # of Vcoils > 0");
  std::string err = lmat_full_host(m, Lmat);
  if (!err.empty()) return fail(err);
  return 0;
}

int thincurr_b200_set_coils(void* tw_ptr, int kind, int nsets, const int* set_ptr, const int* fil_ptr, const double* pts,
                            const double* scales, const double* radius, const double* res_per_len, const int* sens_mask,
                            int* sizes) {
  Model& m = *(Model*)tw_ptr;
  std::vector<CoilSet> sets(nsets);
  for (int s = 0; s < nsets; s++) {
    sets[s].sens_mask = sens_mask ? sens_mask[s] != 0 : false;
    for (int f = set_ptr[s]; f < set_ptr[s + 1]; f++) {
      Filament fl;
      fl.pts.assign(pts + 3 * (size_t)fil_ptr[f], pts + 3 * (size_t)fil_ptr[f + 1]);
      fl.scale = scales ? scales[f] : 1.0;
      fl.radius = radius ? radius[f] : -1.0;
      fl.res_per_len = res_per_len ? res_per_len[f] : -1.0;
      if (kind == 0) {
        if (fl.res_per_len < 0.0) return fail("Invalid resistivity for passive coil");
        if (fl.radius < 1.e-6) return fail("Invalid radius for passive coil");
      } else {
        fl.radius = std::max(1.e-6, fl.radius);
      }
      sets[s].coils.push_back(std::move(fl));
    }
  }
  if (kind == 0) m.vcoils = std::move(sets);
  else m.icoils = std::move(sets);
  m.n_vcoils = (int)m.vcoils.size();
  m.n_icoils = (int)m.icoils.size();
  m.nelems = m.np_active + m.nholes + m.n_vcoils;
  m.have_coil_mutuals = false;
  if (sizes) fill_sizes(m, sizes);
  return 0;
}

int thincurr_b200_set_sensors(void* tw_ptr, int nsensors, const int* fil_ptr, const double* pts, const double* scale_fac,
                              void** sensor_ptr) {
  (void)tw_ptr;
  auto* s = new Sensors();
  for (int i = 0; i < nsensors; i++) {
    FluxLoop fl;
    fl.pts.assign(pts + 3 * (size_t)fil_ptr[i], pts + 3 * (size_t)fil_ptr[i + 1]);
    fl.scale_fac = scale_fac ? scale_fac[i] : 1.0;
    char nm[40];
    std::snprintf(nm, sizeof nm, "FLOOP_%d", i);
    fl.name = nm;
    s->floops.push_back(std::move(fl));
  }
  *sensor_ptr = s;
  return 0;
}

int thincurr_b200_msensor(void* tw_ptr, void* sensor_ptr, void** Ms_ptr, void** Msc_ptr) {
  Model& m = *(Model*)tw_ptr;
  std::string err = gpu_msensor(m, *(Sensors*)sensor_ptr);
  if (!err.empty()) return fail(err);
  *Ms_ptr = m.Ael2sen.p;
  *Msc_ptr = m.Adr2sen.p;
  return 0;
}

int thincurr_b200_plan(void* tw_ptr, int nshards, int shard, int* nrows) {
  Model& m = *(Model*)tw_ptr;
  std::string err = ensure_plan(m, nshards);
  if (!err.empty()) return fail(err);
  if (nshards < 1 || shard < 0 || shard >= nshards) return fail("Invalid shard index");
  int p0, p1;
  std::vector<int> rows;
  shard_rows(m, nshards, shard, p0, p1, rows);
  *nrows = (int)rows.size();
  return 0;
}

int thincurr_b200_stream_plan(void* tw_ptr, int* nbands, int* band_ref_ptr, int* band_patch_ptr) {
  Model& m = *(Model*)tw_ptr;
  std::string err;
  *nbands = 0;
  if (!ensure_banded_plan(m, err)) return err.empty() ? 0 : fail(err);
  const Plan& pl = *m.plan;
  *nbands = (int)pl.band_ref_ptr.size() - 1;
  for (size_t b = 0; b < pl.band_ref_ptr.size(); b++) {
    if (band_ref_ptr) band_ref_ptr[b] = pl.band_ref_ptr[b];
    if (band_patch_ptr) band_patch_ptr[b] = pl.band_patch_ptr[b];
  }
  return 0;
}

int thincurr_b200_dof_patches(void* tw_ptr, int nshards, int* patch_of_dof) {
  // patch index of every vertex / hole DOF (reference id) in the plan made for `nshards` shards: with
  // thincurr_b200_shard_rows_sym this tells which entries of a symmetric shard were evaluated in place -- tile {pa,pb}
  // between two shards belongs to the rows of pa iff ((pa + pb) even) == (pa < pb)
  Model& m = *(Model*)tw_ptr;
  std::string err = ensure_plan(m, nshards);
  if (!err.empty()) return fail(err);
  const PatchSet& ps = m.plan->ps;
  for (int p = 0; p < ps.npatch; p++)
    for (int i = ps.patch_dof_ptr[p]; i < ps.patch_dof_ptr[p + 1]; i++) patch_of_dof[ps.dof_orig[i]] = p;
  return 0;
}

int thincurr_b200_plan_chunks(void* tw_ptr, int* patch_chunk_ptr, double* chunk_info) {
  // introspection: chunk_info[nchunk][6] = centre (3), radius, longest edge, cells; patch_chunk_ptr[npatch+1]
  Model& m = *(Model*)tw_ptr;
  std::string err = ensure_plan(m);
  if (!err.empty()) return fail(err);
  const PatchSet& ps = m.plan->ps;
  if (patch_chunk_ptr) std::memcpy(patch_chunk_ptr, ps.patch_chunk_ptr.data(), ps.patch_chunk_ptr.size() * sizeof(int));
  if (chunk_info)
    for (int c = 0; c < ps.nchunk; c++) {
      const ChunkMeta& cm = ps.chunks[c];
      double* o = chunk_info + 6 * (size_t)c;
      o[0] = cm.cx; o[1] = cm.cy; o[2] = cm.cz; o[3] = cm.rad; o[4] = cm.emax; o[5] = cm.ncell;
    }
  return 0;
}

int64_t thincurr_b200_model_bytes(void* tw_ptr) {
  Model& m = *(Model*)tw_ptr;
  return m.dev.empty() ? 0 : (int64_t)m.dev[0]->ps.bytes;
}

int thincurr_b200_plan_info(void* tw_ptr, int64_t* info) {
  Model& m = *(Model*)tw_ptr;
  std::string err = ensure_plan(m);
  if (!err.empty()) return fail(err);
  const PatchSet& ps = m.plan->ps;
  std::vector<Tile> tiles;
  build_self_tiles(ps, 0, ps.npatch, tiles);
  int64_t cells = 0, chunk_pairs = 0, cell_pairs = 0;
  for (int n : ps.patch_ncell) cells += n;
  for (auto& t : tiles) {
    int64_t na = ps.patch_chunk_ptr[t.pa + 1] - ps.patch_chunk_ptr[t.pa], nb = ps.patch_chunk_ptr[t.pb + 1] - ps.patch_chunk_ptr[t.pb];
    chunk_pairs += na * nb;
    cell_pairs += (int64_t)ps.patch_ncell[t.pa] * ps.patch_ncell[t.pb];
  }
  info[0] = m.plan->patch_size;
  info[1] = ps.npatch;
  info[2] = ps.nchunk;
  info[3] = cells;
  info[4] = (int64_t)tiles.size();
  info[5] = chunk_pairs;
  info[6] = cell_pairs;
  info[7] = ps.nvert_patch;
  return 0;
}

int thincurr_b200_shard_rows(void* tw_ptr, int nshards, int shard, int* row_ids) {
  Model& m = *(Model*)tw_ptr;
  std::string err = ensure_plan(m, nshards);
  if (!err.empty()) return fail(err);
  int p0, p1;
  std::vector<int> rows;
  shard_rows(m, nshards, shard, p0, p1, rows);
  std::copy(rows.begin(), rows.end(), row_ids);
  return 0;
}

int thincurr_b200_Lmat_shard(void* tw_ptr, int nshards, int shard, double* d_out, int64_t ld, void* stream, int64_t* stats) {
  Model& m = *(Model*)tw_ptr;
  unsigned long long st[8] = {0};
  std::string err = lmat_shard_device(m, nshards, shard, d_out, ld, (cudaStream_t)stream, stats ? st : nullptr);
  if (!err.empty()) return fail(err);
  if (stats)
    for (int k = 0; k < 8; k++) stats[k] = (int64_t)st[k];
  return 0;
}

int thincurr_b200_shard_rows_sym(void* tw_ptr, int nshards, int shard, int* nrows, int* row_ids) {
  Model& m = *(Model*)tw_ptr;
  std::string err = ensure_plan(m, nshards);
  if (!err.empty()) return fail(err);
  int p0, p1;
  std::vector<int> rows;
  shard_rows(m, nshards, shard, p0, p1, rows, true);
  if (nrows) *nrows = (int)rows.size();
  if (row_ids) std::copy(rows.begin(), rows.end(), row_ids);
  return 0;
}

int thincurr_b200_Lmat_shard_sym(void* tw_ptr, int nshards, int shard, double* d_out, int64_t ld, void* stream, int64_t* stats) {
  Model& m = *(Model*)tw_ptr;
  unsigned long long st[8] = {0};
  std::string err = lmat_shard_device(m, nshards, shard, d_out, ld, (cudaStream_t)stream, stats ? st : nullptr, true);
  if (!err.empty()) return fail(err);
  if (stats)
    for (int k = 0; k < 8; k++) stats[k] = (int64_t)st[k];
  return 0;
}

int thincurr_b200_Lmat_shard_host(void* tw_ptr, int nshards, int shard, double* h_out, int64_t ld, int64_t* stats) {
  // end-to-end: upload the model, build the rows band by band, bring every finished band back to host memory
  // (copy engine, overlapped with the evaluation of the next band)
  Model& m = *(Model*)tw_ptr;
  const bool trace = std::getenv("THINCURR_B200_TRACE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  auto t0 = now();
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return fail("No CUDA device available (there is no CPU fallback)");
  std::string err = ensure_plan(m, nshards);
  if (!err.empty()) return fail(err);
  std::shared_ptr<DeviceState> ds;
  err = ensure_device(m, device, ds);
  if (!err.empty()) return fail(err);
  err = ds->ps.upload_from(m.plan->ps);  // the step's inputs: host model -> device, every call
  if (!err.empty()) return fail(err);
  auto t1 = now();
  int p0, p1;
  std::vector<int> rows;
  shard_rows(m, nshards, shard, p0, p1, rows);
  double* d = nullptr;
  const size_t bytes = rows.size() * (size_t)ld * 8;
  err = scratch_rows(*ds, bytes, &d);
  if (!err.empty()) return fail(err);
  auto t2 = now();
  cudaStream_t sb = nullptr, sc = nullptr;
  if (cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking) != cudaSuccess)
    return fail("cudaStreamCreate failed");
  const bool pinned = is_pinned(h_out);
  HostCopier copier(device);
  std::vector<cudaEvent_t> evs;
  auto band_done = [&](int, int r0, int r1) -> std::string {
    cudaEvent_t ev = nullptr;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return "cudaEventCreate failed";
    evs.push_back(ev);
    if (cudaEventRecord(ev, sb) != cudaSuccess) return "cudaEventRecord failed";
    const size_t off = (size_t)r0 * ld, nb = (size_t)(r1 - r0) * ld * 8;
    if (!pinned) {
      copier.push(h_out + off, d + off, nb, ev);
      return "";
    }
    if (cudaStreamWaitEvent(sc, ev, 0) != cudaSuccess || cudaMemcpyAsync(h_out + off, d + off, nb, cudaMemcpyDeviceToHost, sc) != cudaSuccess)
      return std::string("Device->host copy failed: ") + cudaGetErrorString(cudaGetLastError());
    return "";
  };
  // (device-side evaluation counters would force a stream synchronisation per band: not collected on this path)
  const int nb = m.n_vcoils == 0 ? auto_bands(m.plan->ps, p0, p1) : 1;
  if (m.n_vcoils == 0) {
    err = lmat_shard_device(m, nshards, shard, d, ld, sb, nullptr, false, nb, band_done);
  } else {  // V-coil rows and columns are filled after the tiles: one copy of the finished block
    err = lmat_shard_device(m, nshards, shard, d, ld, sb, nullptr, false);
    if (err.empty()) err = band_done(0, 0, (int)rows.size());
  }
  auto t3 = now();
  if (cudaStreamSynchronize(sb) != cudaSuccess && err.empty()) err = std::string("Kernel failed: ") + cudaGetErrorString(cudaGetLastError());
  if (cudaStreamSynchronize(sc) != cudaSuccess && err.empty()) err = std::string("Device->host copy failed: ") + cudaGetErrorString(cudaGetLastError());
  {
    std::string ce = copier.finish();
    if (err.empty()) err = ce;
  }
  auto t4 = now();
  for (cudaEvent_t e : evs) cudaEventDestroy(e);
  cudaStreamDestroy(sb);
  cudaStreamDestroy(sc);
  if (trace)
    std::fprintf(stderr, "[Lmat_shard_host] plan+upload %.1f ms, scratch %.1f, enqueue %.1f (%d bands), drain %.1f (%.2f GB, %s)\n", ms(t0, t1),
                 ms(t1, t2), ms(t2, t3), nb, ms(t3, t4), bytes * 1e-9, pinned ? "pinned" : "pageable");
  if (!err.empty()) return fail(err);
  if (stats) {
    for (int k = 0; k < 8; k++) stats[k] = 0;
    stats[4] = nb;
    stats[5] = (int64_t)ds->ps.bytes;  // host->device bytes of the model upload
    stats[6] = (int64_t)bytes;         // device->host bytes
  }
  return 0;
}

int thincurr_b200_release_device(void* tw_ptr) {
  if (!tw_ptr) return 0;
  drop_device_state(*(Model*)tw_ptr);
  return 0;
}

int thincurr_b200_Lmat_block(void* tw_ptr, int nrows, const int* row_ids, int ncols, const int* col_ids, double* d_out, int64_t ld,
                             void* stream_) {
  // dense block L(row_ids, col_ids) of the self-inductance matrix (the evaluator behind tw_compute_Lmatblock,
  // thin_wall_hodlr.F90:136-404: same pair integrals and role rule restricted to a row / column DOF subset)
  Model& m = *(Model*)tw_ptr;
  cudaStream_t stream = (cudaStream_t)stream_;
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return fail("No CUDA device available (there is no CPU fallback)");
  std::shared_ptr<DeviceState> ds;
  std::string err = ensure_device(m, device, ds);
  if (!err.empty()) return fail(err);
  const PatchSet& ps = m.plan->ps;
  if (ld < ncols) return fail("thincurr_b200_Lmat_block: ld < ncols");
  std::vector<int> internal(ps.ndof, -1), patch_of(ps.ndof, 0);
  for (int i = 0; i < ps.ndof; i++) internal[ps.dof_orig[i]] = i;
  for (int p = 0; p < ps.npatch; p++)
    for (int i = ps.patch_dof_ptr[p]; i < ps.patch_dof_ptr[p + 1]; i++) patch_of[i] = p;
  std::vector<int> row_out(ps.ndof, -1), col_map(ps.ndof, -1);
  std::vector<char> prow(ps.npatch, 0), pcol(ps.npatch, 0);
  for (int r = 0; r < nrows; r++) {
    if (row_ids[r] < 0 || row_ids[r] >= ps.ndof) return fail("thincurr_b200_Lmat_block: row id out of range (vertex and hole DOFs only)");
    const int i = internal[row_ids[r]];
    if (row_out[i] >= 0) return fail("thincurr_b200_Lmat_block: duplicate row id");
    row_out[i] = r;
    prow[patch_of[i]] = 1;
  }
  for (int c = 0; c < ncols; c++) {
    if (col_ids[c] < 0 || col_ids[c] >= ps.ndof) return fail("thincurr_b200_Lmat_block: column id out of range (vertex and hole DOFs only)");
    if (col_map[col_ids[c]] >= 0) return fail("thincurr_b200_Lmat_block: duplicate column id");
    col_map[col_ids[c]] = c;
    pcol[patch_of[internal[col_ids[c]]]] = 1;
  }
  // every (row patch, column patch) pair as an ordinary tile: all entries, both roles where needed, no mirror
  std::vector<Tile> tiles;
  for (int pa = 0; pa < ps.npatch; pa++)
    if (prow[pa])
      for (int pb = 0; pb < ps.npatch; pb++)
        if (pcol[pb]) tiles.push_back(Tile{pa, pb, 4, (float)ps.patch_ncell[pa] * ps.patch_ncell[pb]});
  std::stable_sort(tiles.begin(), tiles.end(), [](const Tile& a, const Tile& b) { return a.cost > b.cost; });
  if (cudaMemset2DAsync(d_out, (size_t)ld * 8, 0, (size_t)ncols * 8, (size_t)nrows, stream) != cudaSuccess)
    return fail("cudaMemset2DAsync failed on the output block");
  int* d_col_map = nullptr;
  if (cudaMallocAsync((void**)&d_col_map, (size_t)ps.ndof * sizeof(int), stream) != cudaSuccess) return fail("Device allocation failed");
  cudaMemcpyAsync(d_col_map, col_map.data(), (size_t)ps.ndof * sizeof(int), cudaMemcpyHostToDevice, stream);
  cudaStreamSynchronize(stream);  // col_map is a pageable host vector
  err = gpu_lmat_tiles(ds->ps, ds->ps, tiles, row_out, true, d_out, ld, stream, nullptr, d_col_map, false);
  cudaFreeAsync(d_col_map, stream);
  if (!err.empty()) return fail(err);
  return 0;
}

int thincurr_b200_Lmatblock(void* tw_row, void* tw_col, int nrp, const int* row_pts, int ncp, const int* col_pts, double* out, int64_t ld,
                            void* stream) {
  if (!tw_row) return fail("thincurr_b200_Lmatblock: no model");
  Model &mr = *(Model*)tw_row, &mc = tw_col ? *(Model*)tw_col : mr;
  std::string err = gpu_lmatblock(mr, mc, nrp, row_pts, ncp, col_pts, out, ld, (cudaStream_t)stream);
  if (!err.empty()) return fail(err);
  return 0;
}

int thincurr_b200_LmatHole(void* tw_ptr, double* out, int64_t ld, void* stream) {
  std::string err = gpu_lmathole(*(Model*)tw_ptr, out, ld, (cudaStream_t)stream);
  if (!err.empty()) return fail(err);
  return 0;
}

int thincurr_b200_Bops_block(void* tw_ptr, int nrp, const int* row_pts, int ncp, const int* col_pts, int dir, double* out, int64_t ld,
                             void* stream) {
  std::string err = gpu_bops_block(*(Model*)tw_ptr, nrp, row_pts, ncp, col_pts, dir, out, ld, (cudaStream_t)stream);
  if (!err.empty()) return fail(err);
  return 0;
}

int thincurr_b200_cross_eval(void* tw_ptr1, void* tw_ptr2, int nrhs, const double* vec1, double* vec2, int64_t* counts) {
  long long c[3] = {0, 0, 0};
  std::string err = gpu_cross_eval(*(Model*)tw_ptr1, *(Model*)tw_ptr2, nrhs, vec1, vec2, counts ? c : nullptr);
  if (!err.empty()) return fail(err);
  if (counts)
    for (int k = 0; k < 3; k++) counts[k] = c[k];
  return 0;
}

int thincurr_b200_h5_write(const char* path, int nitems, const char* const* names, const int* is_f64, const int* ranks,
                           const int64_t* dims, const void* const* data) {
  std::vector<H5Item> items;
  size_t o = 0;
  for (int i = 0; i < nitems; i++) {
    H5Item it;
    it.name = names[i];
    it.f64 = is_f64[i] != 0;
    for (int k = 0; k < ranks[i]; k++) it.dims.push_back((uint64_t)dims[o++]);
    it.data = data[i];
    items.push_back(it);
  }
  std::string err = write_h5_file(path, items);
  if (!err.empty()) return fail(err);
  return 0;
}

int thincurr_b200_Bel_shard(void* tw_ptr, int nshards, int shard, double* d_out, void* stream) {
  Model& m = *(Model*)tw_ptr;
  std::string err = bel_shard_device(m, nshards, shard, d_out, (cudaStream_t)stream);
  if (!err.empty()) return fail(err);
  return 0;
}

int thincurr_b200_pair_stats(void* tw_ptr, int64_t* hist, int64_t* visited) {
  Model& m = *(Model*)tw_ptr;
  std::string err = gpu_pair_stats(m, hist, visited);
  if (!err.empty()) return fail(err);
  return 0;
}

long long thincurr_b200_launch_count(void) { return launch_count(); }

double thincurr_b200_dfma_peak(int device, double* sm_clock_mhz) { return gpu_dfma_peak(device, sm_clock_mhz); }

int thincurr_b200_get_model(void* tw_ptr, int* pmap, int* lc, int* kfh, int* lfh, double* qbasis, double* ca) {
  Model& m = *(Model*)tw_ptr;
  if (pmap) std::copy(m.pmap.begin(), m.pmap.end(), pmap);
  if (lc) std::copy(m.lc.begin(), m.lc.end(), lc);
  if (kfh) std::copy(m.kfh.begin(), m.kfh.end(), kfh);
  if (lfh) std::copy(m.lfh.begin(), m.lfh.end(), lfh);
  if (qbasis) std::copy(m.qbasis.begin(), m.qbasis.end(), qbasis);
  if (ca) std::copy(m.ca.begin(), m.ca.end(), ca);
  return m.nfh;
}

int thincurr_b200_hashes(void* tw_ptr, int32_t* hash_lc, int32_t* hash_r) {
  Model& m = *(Model*)tw_ptr;
  *hash_lc = m.hash_lc();
  *hash_r = m.hash_r();
  return 0;
}

}  // extern "C"
