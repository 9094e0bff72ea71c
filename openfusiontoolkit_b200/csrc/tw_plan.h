// tw_plan.h -- owner-computes decomposition of the dense operator builds.
//
// DOFs (vertex + hole DOFs of thin_wall.F90:282-342) are grouped into spatially compact
// PATCHES by recursive coordinate bisection; a patch owns its rows of L.  Every patch carries
// the list of cells that touch any of its DOFs (a one-ring halo), cut into CHUNKS of <= kCH
// cells whose geometry is stored SoA and contiguous in HBM so a CTA can stage a chunk in shared
// memory with bulk copies.  An output TILE = (row patch, column patch) is owned by exactly one
// CTA: it evaluates the pair integrals T(c1,c2) chunk pair by chunk pair into a shared-memory
// tile and contracts them onto the 3x3 (+hole) vertex DOFs, so no atomics are ever needed.
#pragma once
#include <cstdint>
#include <vector>

#include "tw_host.h"

namespace tw {

constexpr int kCH = 64;        // cells per chunk
constexpr int kMaxChunkDof = 3 * kCH;   // DOFs with incidences in one chunk
constexpr int kMaxChunkInc = 4 * kCH;   // incidences in one chunk
constexpr int kGeomRows = 25;  // doubles per cell in the SoA chunk record: 0-8 vertices, 9 area, 10-18 qbasis,
                               // 19-21 unit normal as tw_compute_phipot forms it, 22-24 mesh normal (trimesh_norm)

constexpr int kRowHalf = 64;   // row cells per pass of the L kernel (a chunk is swept in kCH/kRowHalf passes)

// Per-chunk index record the L kernel stages with one bulk copy (size is a multiple of 16 bytes).
struct alignas(16) ChunkAux {
  int dmin[kCH], dmax[kCH];          // min / max reference DOF id per cell (in-patch incidences)
  int orig[kMaxChunkDof];            // reference DOF id per local DOF
  int iptr[kMaxChunkDof + 4];        // CSR offsets of the incidence list (ndof+1 used)
  uint16_t inc[kMaxChunkInc];        // cell(6b) | local vertex(2b)<<6 | negative<<8
  int nact[kCH / kRowHalf];          // local DOFs that have a cell in row half h ...
  unsigned char act[kCH / kRowHalf][kMaxChunkDof];  // ... and their list
};
static_assert(sizeof(ChunkAux) % 16 == 0, "ChunkAux must be a multiple of 16 bytes (bulk copy)");

struct ChunkMeta {
  int ncell;    // valid cells in the chunk
  int ndof;     // DOFs that have an incidence in this chunk
  int dof_off;  // offset of this chunk's DOF list (chunk_dof / inc_ptr(+chunk id))
  int inc_off;  // offset of this chunk's incidence list
  double cx, cy, cz;  // centre of the bounding box of the chunk's vertices
  double rad;         // radius of the bounding sphere around that centre
  double emax;        // longest triangle edge of the chunk's cells (bounds dl_max - dl_min of a cell pair by emax_I + emax_J)
};

// Patch decomposition of ONE model (row or column side).
struct PatchSet {
  int npatch = 0, nvert_patch = 0, nchunk = 0, ndof = 0;  // ndof = np_active + nholes
  std::vector<int> patch_dof_ptr;    // [npatch+1] internal DOF index range per patch
  std::vector<int> dof_orig;         // [ndof] internal index -> reference DOF id (0-based)
  std::vector<int> patch_chunk_ptr;  // [npatch+1]
  std::vector<int> patch_ncell;      // [npatch]
  std::vector<ChunkMeta> chunks;     // [nchunk]
  std::vector<double> geom;          // [nchunk][kGeomRows][kCH]
  std::vector<int> cell_dmin, cell_dmax;  // [nchunk][kCH] min/max reference DOF id (in-patch incidences)
  std::vector<int> chunk_dof;        // internal DOF index per (chunk, local dof)
  std::vector<int> chunk_inc_ptr;    // per chunk: ndof+1 offsets (relative to inc_off), stored at dof_off+chunk
  std::vector<uint16_t> inc;         // cell(6b) | local vertex(2b)<<6 | negative<<8
  std::vector<int> cell_ids;         // [nchunk][kCH] reference cell id (debug / stats)
  mutable std::vector<float> self_cost;  // [npatch][npatch] cost estimate of the self tiles (filled on first use)
};

struct Tile {
  int pa, pb;   // row patch, column patch
  int flags;    // bit0: diagonal (pa==pb, only entries a<=b computed, mirrored)
                // bit1: mirror-write the transposed entries (pb rows are owned too)
                // bit2: second role (T with the column cell analytic) may be needed
                // bit3: the transposed entries are not written by the tile kernel (rows of both patches are in
                //       the output block: one symmetrisation pass copies them afterwards)
  float cost;
};

struct Plan {
  int patch_size = 0;
  int nshards_hint = 1;  // shard count the patch size was chosen for
  int serial = 0;        // distinguishes successive plans of a model (device mirrors are re-uploaded when it changes)
  // banded plan (streamed single-device build, tw_capi.cu): band b = reference DOF ids [band_ref_ptr[b], band_ref_ptr[b+1])
  // = patches [band_patch_ptr[b], band_patch_ptr[b+1]) (the hole patches belong to the last band); empty otherwise
  std::vector<int> band_ref_ptr, band_patch_ptr;
  int band_ndev = 0;     // devices the bands were sized for
  PatchSet ps;
};

// Patch size (DOFs) chosen for a model with nv vertex DOFs whose rows are spread over nshards pieces of work
int auto_patch_size(int nv, int nshards);
// Build the patch decomposition of a model; P = target DOFs per patch (0 = choose automatically).
// ref_cuts (optional, ascending, first 0, last np_active): the vertex DOFs are first cut into these ranges of reference ids
// and every range gets patches of its own, in range order (band_patch_ptr: first vertex patch of every range + total)
std::string build_patches(const Model& m, int P, PatchSet& out, int nshards = 1, const std::vector<int>* ref_cuts = nullptr,
                          std::vector<int>* band_patch_ptr = nullptr);
// Unit normal of a triangle exactly as tw_compute_phipot evaluates it (thin_wall.F90:1942-1943): IEEE operations
// in the reference's order, no contraction (the device reads these values instead of recomputing them).
void phipot_normal(const double* P /*[3][3]*/, double* n);
// Contiguous patch range [p0,p1) of shard `shard` out of `nshards`, balanced by pair count
void shard_range(const PatchSet& ps, int nshards, int shard, int& p0, int& p1);
// Same for the symmetric build: every pair integral is evaluated on exactly one shard.  Tiles inside a shard's diagonal
// block are its own; a tile {pa,pb} between two shards is evaluated by the shard owning `pa` iff sym_tile_is_mine(pa,pb)
// (a checkerboard over the patch pairs: every off-diagonal block is split evenly between its two shards, so all shards
// do the same work, hold the same number of rows and exchange the same volume); the other shard copies the transposed
// block afterwards (one exchange after the assembly).
void shard_range_sym(const PatchSet& ps, int nshards, int shard, int& p0, int& p1);
#ifdef __CUDACC__
__host__ __device__
#endif
inline bool sym_tile_is_mine(int pa /*my row patch*/, int pb /*patch of another shard*/) {
  const int lo = pa < pb ? pa : pb;
  return (((pa + pb) & 1) == 0) == (lo == pa);
}
// Tiles of a self-inductance build for row patches [p0,p1)
// upper_only: symmetric shards -- of the tiles against unowned patches only those sym_tile_is_mine assigns to these rows
// skip_lo >= 0: also skip the tiles against patches [skip_lo, p0) -- the rows of earlier bands of the same build, whose
// transposed blocks are copied into place afterwards (banded builds, tw_capi.cu)
void build_self_tiles(const PatchSet& ps, int p0, int p1, std::vector<Tile>& tiles, bool upper_only = false, int skip_lo = -1);
// Tiles of a mutual build (all row patches x all column patches)
void build_mutual_tiles(const PatchSet& rows, const PatchSet& cols, std::vector<Tile>& tiles);

}  // namespace tw
