// tw_host.h -- host-side ThinCurr model for the B200 operator builds.
//
// Mirrors the state the reference keeps in `tw_type` (src/physics/thin_wall.F90:111-154)
// but only what the dense operator builds read: oriented mesh, areas, P1 surface-curl
// basis (`qbasis`), DOF map (`pmap`), hole incidence CSR (`kfh/lfh`), coil filaments,
// sensors.  Everything here is O(N) CPU work; the O(N^2) builds live in tw_kernels.cu.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace tw {

constexpr double kPi = 3.14159265358979323846;
constexpr double kMu0 = kPi * 4.e-7;  // src/base/oft_local.F90 mu0

struct Filament {
  std::vector<double> pts;  // [npts][3]
  double scale = 1.0, radius = -1.0, res_per_len = -1.0;
  int npts() const { return (int)(pts.size() / 3); }
};
struct CoilSet {
  std::vector<Filament> coils;
  bool sens_mask = false;
  std::string name;
  double Lself = 0.0, Rself = 0.0;
};
struct FluxLoop {
  std::vector<double> pts;  // [np][3]
  double scale_fac = 1.0;
  std::string name;
};
struct Sensors {
  std::vector<FluxLoop> floops;
  int njumpers = 0;
};

// Minimal XML element tree (only what <thincurr> input needs).
struct XmlNode {
  std::string tag, text;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<XmlNode>> kids;
  const XmlNode* child(const std::string& t) const;
  std::vector<const XmlNode*> children(const std::string& t) const;
};
std::unique_ptr<XmlNode> xml_parse_file(const std::string& path, std::string& err);

// Flattened filament lists as uploaded to the device.
struct FlatCoils {
  std::vector<int> set_ptr{0}, fil_ptr{0}, sens_mask;
  std::vector<double> pts, scales, radius;
  int nsets() const { return (int)set_ptr.size() - 1; }
  int nfil() const { return (int)fil_ptr.size() - 1; }
  int npts() const { return (int)(pts.size() / 3); }
  void append(const CoilSet& s);
};

struct DeviceState;  // defined in tw_kernels.cu (per-device mirrors + plan)
struct Plan;         // defined in tw_plan.cpp

struct HostBuf {  // pinned host buffer owned by the model
  double* p = nullptr;
  size_t n = 0;
  bool pinned = false;
  void alloc(size_t count, bool zero = true);
  void release();
  ~HostBuf() { release(); }
};

struct Model {
  // ---- mesh (0-based; lc after orientation sync) ----
  int np = 0, nc = 0, ne = 0, nreg = 1;
  std::vector<double> r;       // [np][3]
  std::vector<int> lc;         // [nc][3]
  std::vector<int> reg;        // [nc] 1-based
  std::vector<double> ca, va;  // cell / vertex areas
  std::vector<double> norm;    // [nc][3]
  std::vector<double> qbasis;  // [nc][3 vert][3 xyz]
  // linkage
  std::vector<int> le;         // [ne][2] (lo,hi)
  std::vector<int> lce;        // [nc][3] edge id per local slot (slot j opposite-ish, tri_ed)
  std::vector<int> lcc;        // [nc][3] neighbour cell or -1
  std::vector<int> kec, lec;   // edge -> cells (ascending cell id)
  std::vector<int> kpc, lpc;   // point -> cells (ascending cell id)
  std::vector<int> kpe, lpe;   // point -> edges (ascending edge id)
  std::vector<char> be, bp;    // boundary edge / point flags
  int nflipped = 0;
  bool keep_orientation = false;  // skip the orientation sync (cells come from a host that already ran it)
  // ---- DOFs ----
  int np_active = 0, nholes = 0, n_vcoils = 0, n_icoils = 0, nelems = 0, nfh = 0;
  std::vector<int> pmap;       // [np] 1-based DOF id, 0 = inactive
  std::vector<int> kfh;        // [nc+1]
  std::vector<int> lfh;        // [nfh][2] (signed 1-based hole id, 0-based local vertex)
  std::vector<std::vector<int>> hole_chain;
  std::vector<int> closures;   // closure vertices (0-based)
  int n_jumper_sets = 0;       // nodesets from `jumper_start` on (thincurr_f.F90:172-190): kept as a count only
  // ---- physics inputs ----
  std::vector<double> eta_surf, eta_vol, thickness;  // eta stored as eta/mu0 (thin_wall.F90:2864)
  std::vector<int> sens_mask;                         // [nreg]
  std::vector<CoilSet> vcoils, icoils;
  // ---- operators (library-owned, reference column-major layouts) ----
  HostBuf Lmat, Ael2coil, Ael2dr, Acoil2coil, Ael2sen, Adr2sen, Bel, Bdr;
  bool have_coil_mutuals = false;
  int nsensors_built = 0;
  std::vector<int> R_kr, R_lc;  // 1-based CSR
  std::vector<double> R_val;
  // ---- device side ----
  std::shared_ptr<Plan> plan;
  bool no_stream_plan = false;  // the banded plan of the streamed build was tried and rejected (reference numbering without locality)
  std::vector<std::shared_ptr<DeviceState>> dev;  // one per CUDA device used
  std::shared_ptr<void> block_ctx;                // plain cell arrays on the device for the list sweeps (tw_blocks.cu)
  bool verbose = true;

  // setup (tw_setup.cpp)
  std::string setup_from_arrays(int np_, const double* r_, int nc_, const int* lc1, const int* reg_,
                                const int* pmap_in, const std::vector<std::vector<int>>& nodesets0,
                                const std::vector<int>& closure_cells0, const XmlNode* thincurr_xml);
  // model from the arrays a Fortran host already holds in tw_type (oriented lc, pmap, hole CSR):
  // no orientation sync, no hole / DOF construction (thin_wall.F90:111-154)
  std::string setup_from_tw(int np_, const double* r_, int nc_, const int* lc1, const int* reg_, const int* pmap1,
                            int np_active_, int nholes_, const int* kfh1, const int* lfh1, const double* ca_,
                            const double* qbasis_);
  std::string load_coils_xml(const XmlNode* group, const char* prefix, std::vector<CoilSet>& out);
  std::string load_eta_xml(const XmlNode* tc);
  void build_rmat();
  int32_t hash_lc() const;
  int32_t hash_r() const;

 private:
  std::string mesh_init();
  void sync_face_normals();
  void invert_cell(int c);
  std::string build_holes(const std::vector<std::vector<int>>& nodesets0);
  std::string hole_pseq(int i0, std::vector<int>& chain);
  std::string order_hole_list(const std::vector<int>& in, std::vector<int>& out);
  std::string setup_hole(const std::vector<int>& lp, std::vector<int>& cells_signed, std::vector<int>& kpc_h);
  std::string build_pmap(const int* pmap_in, const std::vector<int>& closure_cells0);
  void geometry();
  void cell_normal(int c, double* n) const;
  int find_edge(int a, int b) const;
};

// native mesh file (HDF5 superblock v0) and sensor file readers
struct NativeMesh {
  int np = 0, nc = 0;
  std::vector<double> r;  // [np][3]
  std::vector<int> lc;    // [nc][3] 1-based
  std::vector<int> reg, pmap;
  std::vector<std::vector<int>> nodesets, sidesets;  // 1-based
};
std::string read_native_mesh(const std::string& path, NativeMesh& out);
// minimal HDF5 writer / raw readers (operator caches, thin_wall.F90:2175-2225): root-level contiguous datasets
struct H5Item {
  std::string name;
  bool f64;                    // float64, else int32
  std::vector<uint64_t> dims;  // C order (slowest first)
  const void* data;
};
std::string write_h5_file(const std::string& path, const std::vector<H5Item>& items);
std::string read_h5_dataset_f64_into(const std::string& path, const std::string& name, double* dst, uint64_t count);
std::string read_h5_dataset_i32(const std::string& path, const std::string& name, std::vector<int32_t>& out);
std::string read_h5_dataset_f64(const std::string& path, const std::string& name, std::vector<double>& out,
                                std::vector<uint64_t>& shape);
std::string read_floops(const std::string& path, Sensors& out);
int32_t simple_hash(const void* key, long length);

// Fortran unformatted sequential records (gfortran framing), used by the operator caches
bool funf_write_record(FILE* f, const void* data, size_t bytes);
bool funf_read_record(FILE* f, void* data, size_t bytes);

}  // namespace tw
