// tw_capi.cu -- C ABI of libthincurr_b200.so (see include/thincurr_b200.h).
//
// Block 1 mirrors the BIND(C) wrappers of the reference (src/python/wrappers/thincurr_f.F90,
// oft_base_f.F90) for the operator-build path; block 2 is the sharded device interface.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
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

void HostBuf::alloc(size_t count, bool zero) {
  if (p && n == count) {  // reuse (rebuilds of the same operator keep the Python view valid)
    if (zero) std::memset(p, 0, count * sizeof(double));
    return;
  }
  release();
  n = count;
  if (count == 0) return;
  size_t bytes = count * sizeof(double);
  // pinned memory for matrices up to 16 GiB; beyond that pinning costs more than it saves
  if (bytes <= (size_t)16 << 30) {
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

std::string ensure_plan(Model& m) {
  if (m.plan) return "";
  auto pl = std::make_shared<Plan>();
  int P = 0;
  if (const char* e = std::getenv("THINCURR_B200_PATCH")) P = std::atoi(e);
  std::string err = build_patches(m, P, pl->ps);
  if (!err.empty()) return err;
  m.plan = pl;
  return "";
}

std::string ensure_device(Model& m, int device, std::shared_ptr<DeviceState>& out) {
  std::string err = ensure_plan(m);
  if (!err.empty()) return err;
  for (auto& d : m.dev)
    if (d->device == device) {
      out = d;
      return "";
    }
  if (cudaSetDevice(device) != cudaSuccess) return std::string("cudaSetDevice failed: ") + cudaGetErrorString(cudaGetLastError());
  auto ds = std::make_shared<DeviceState>();
  ds->device = device;
  err = ds->ps.upload_from(m.plan->ps);
  if (!err.empty()) return err;
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

// Build the self-inductance rows of one shard on the current device into d_out[nrows][ld]
// (zeroed here).  Asynchronous on `stream` unless stats are requested.
// sym: symmetric partition, only the blocks against this and later shards are computed (upper trapezoid)
std::string lmat_shard_device(Model& m, int nshards, int shard, double* d_out, long long ld, cudaStream_t stream,
                              unsigned long long* stats, bool sym) {
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return "No CUDA device available (there is no CPU fallback)";
  std::shared_ptr<DeviceState> ds;
  std::string err = ensure_device(m, device, ds);
  if (!err.empty()) return err;
  if (m.n_vcoils > 0 && !m.have_coil_mutuals) return "Coil mutuals required if, # of Vcoils > 0";
  const PatchSet& ps = m.plan->ps;
  int p0, p1;
  std::vector<int> row_ids;
  shard_rows(m, nshards, shard, p0, p1, row_ids, sym);
  std::vector<int> row_out(ps.ndof, -1);
  for (int i = ps.patch_dof_ptr[p0], r = 0; i < ps.patch_dof_ptr[p1]; i++, r++) row_out[i] = r;
  std::vector<Tile> tiles;
  build_self_tiles(ps, p0, p1, tiles, sym);
  if (cudaMemsetAsync(d_out, 0, (size_t)row_ids.size() * ld * sizeof(double), stream) != cudaSuccess)
    return "cudaMemsetAsync failed on the output block";
  err = gpu_lmat_tiles(ds->ps, ds->ps, tiles, row_out, true, d_out, ld, stream, stats);
  if (!err.empty()) return err;
  if (m.n_vcoils > 0) {
    err = gpu_fill_vcoil_block(m, row_ids, d_out, ld, stream);
    if (!err.empty()) return err;
  }
  return "";
}

}  // namespace tw

// =============================================================================================
// Block 1: reference-compatible entry points
// =============================================================================================
extern "C" {

void oftpy_init(int nthreads, bool quiet, const char* input_file, int* slens, void* abort_callback) {
  (void)nthreads; (void)input_file; (void)abort_callback;
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

// full self-inductance matrix into host memory dst[nelems][nelems] (reference layout), rows sharded
// over all visible devices, each shard copied straight to its place.  With several devices that can read each other's
// memory the shards are the symmetric ones (upper trapezoid per device, no pair integral evaluated twice) and every
// device fetches the transposed blocks of the earlier shards over NVLink before its rows go to the host.
static std::string lmat_full_host(Model& m, double* dst) {
  const size_t N = (size_t)m.nelems;
  int ndev = visible_devices();
  if (ndev < 1) return "No CUDA device available (the B200 backend has no CPU fallback)";
  std::string err = ensure_plan(m);
  if (!err.empty()) return err;
  ndev = std::min(ndev, std::max(1, m.plan->ps.npatch));
  bool sym = ndev > 1 && m.n_vcoils == 0 && !std::getenv("THINCURR_B200_FULL_ROWS");
  for (int g = 1; g < ndev && sym; g++)
    for (int s = 0; s < g && sym; s++) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, g, s) != cudaSuccess || !can) sym = false;
    }
  struct Dev {
    double* d = nullptr;
    cudaStream_t s = nullptr;
    cudaEvent_t built = nullptr;
    std::vector<int> rows;
    int i0 = 0, i1 = 0;  // internal DOF range of the rows
    std::shared_ptr<DeviceState> ds;
  };
  std::vector<Dev> devs(ndev);
  const PatchSet& ps = m.plan->ps;
  for (int g = 0; g < ndev && err.empty(); g++) {
    cudaSetDevice(g);
    int p0, p1;
    shard_rows(m, ndev, g, p0, p1, devs[g].rows, sym);
    devs[g].i0 = ps.patch_dof_ptr[p0];
    devs[g].i1 = ps.patch_dof_ptr[p1];
    if (devs[g].rows.empty()) continue;
    if (cudaStreamCreate(&devs[g].s) != cudaSuccess || cudaMalloc((void**)&devs[g].d, devs[g].rows.size() * N * 8) != cudaSuccess) {
      err = std::string("Device allocation failed: ") + cudaGetErrorString(cudaGetLastError());
      break;
    }
    err = lmat_shard_device(m, ndev, g, devs[g].d, (long long)N, devs[g].s, nullptr, sym);
    if (sym && err.empty()) {
      err = ensure_device(m, g, devs[g].ds);
      cudaEventCreateWithFlags(&devs[g].built, cudaEventDisableTiming);
      cudaEventRecord(devs[g].built, devs[g].s);
      for (int s = 0; s < g && err.empty(); s++) {  // transposed blocks of the earlier shards
        if (!devs[s].d) continue;
        cudaError_t pe = cudaDeviceEnablePeerAccess(s, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
          err = std::string("cudaDeviceEnablePeerAccess failed: ") + cudaGetErrorString(pe);
          break;
        }
        cudaGetLastError();
        cudaStreamWaitEvent(devs[g].s, devs[s].built, 0);
        err = gpu_symmetrize_cross(devs[g].ds->ps, devs[g].i0, devs[g].i1, devs[s].i0, devs[s].i1, devs[g].d, devs[s].d, (long long)N, devs[g].s);
      }
    }
    // rows go straight to their place in the reference layout Lmat(:,row)
    for (size_t r = 0; r < devs[g].rows.size() && err.empty();) {
      size_t r1 = r + 1;
      while (r1 < devs[g].rows.size() && devs[g].rows[r1] == devs[g].rows[r1 - 1] + 1) r1++;
      if (cudaMemcpyAsync(dst + (size_t)devs[g].rows[r] * N, devs[g].d + r * N, (r1 - r) * N * 8, cudaMemcpyDeviceToHost,
                          devs[g].s) != cudaSuccess)
        err = std::string("Device->host copy failed: ") + cudaGetErrorString(cudaGetLastError());
      r = r1;
    }
  }
  for (int g = 0; g < ndev; g++) {
    cudaSetDevice(g);
    if (devs[g].s) {
      cudaError_t ce = cudaStreamSynchronize(devs[g].s);
      if (ce != cudaSuccess && err.empty()) err = std::string("Kernel execution failed: ") + cudaGetErrorString(ce);
    }
  }
  for (int g = 0; g < ndev; g++) {  // (all devices are done reading each other's blocks)
    cudaSetDevice(g);
    if (devs[g].s) cudaStreamDestroy(devs[g].s);
    if (devs[g].built) cudaEventDestroy(devs[g].built);
    if (devs[g].d) cudaFree(devs[g].d);
  }
  cudaSetDevice(0);
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
  std::string err = ensure_plan(m);
  if (!err.empty()) return fail(err);
  if (nshards < 1 || shard < 0 || shard >= nshards) return fail("Invalid shard index");
  int p0, p1;
  std::vector<int> rows;
  shard_rows(m, nshards, shard, p0, p1, rows);
  *nrows = (int)rows.size();
  return 0;
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
  std::string err = ensure_plan(m);
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
  std::string err = ensure_plan(m);
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
  // end-to-end: (re)upload the model, build the rows, bring them back to host memory
  Model& m = *(Model*)tw_ptr;
  const bool trace = std::getenv("THINCURR_B200_TRACE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  auto t0 = now();
  drop_device_state(m);
  m.plan.reset();
  std::string err = ensure_plan(m);
  if (!err.empty()) return fail(err);
  auto t1 = now();
  int p0, p1;
  std::vector<int> rows;
  shard_rows(m, nshards, shard, p0, p1, rows);
  double* d = nullptr;
  size_t bytes = rows.size() * (size_t)ld * 8;
  if (cudaMalloc((void**)&d, std::max<size_t>(bytes, 8)) != cudaSuccess) return fail("Device allocation failed");
  auto t2 = now();
  unsigned long long st[8] = {0};
  err = lmat_shard_device(m, nshards, shard, d, ld, 0, stats ? st : nullptr);
  if (err.empty() && cudaDeviceSynchronize() != cudaSuccess) err = std::string("Kernel failed: ") + cudaGetErrorString(cudaGetLastError());
  auto t3 = now();
  if (err.empty() && cudaMemcpy(h_out, d, bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
    err = std::string("Device->host copy failed: ") + cudaGetErrorString(cudaGetLastError());
  auto t4 = now();
  cudaFree(d);
  auto t5 = now();
  if (trace)
    std::fprintf(stderr, "[Lmat_shard_host] plan %.1f ms, alloc %.1f, upload+build %.1f, d2h %.1f (%.2f GB), free %.1f\n", ms(t0, t1), ms(t1, t2),
                 ms(t2, t3), ms(t3, t4), bytes * 1e-9, ms(t4, t5));
  if (!err.empty()) return fail(err);
  if (stats) {
    for (int k = 0; k < 8; k++) stats[k] = (int64_t)st[k];
    stats[5] = m.dev.empty() ? 0 : (int64_t)m.dev[0]->ps.bytes;  // host->device bytes of the model upload
    stats[6] = (int64_t)bytes;                                   // device->host bytes
  }
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
