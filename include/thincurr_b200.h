/* thincurr_b200.h -- C ABI of libthincurr_b200.so
 *
 * Drop-in boundary for ThinCurr's dense operator builds.  The first block re-exports, with
 * identical names / argument lists / ownership and error conventions, the BIND(C) entry
 * points that the reference's Python layer binds from liboftpy.so for this path
 * (reference = OpenFUSIONToolkit @ d08f001b; `F:` = src/python/wrappers/thincurr_f.F90,
 * `P:` = src/python/OpenFUSIONToolkit/ThinCurr/_interface.py).  The second block is the
 * flat ISO_C_BINDING-style interface a Fortran host (thin_wall.F90) or a multi-GPU
 * launcher calls: row-block sharded builds into caller-provided device memory.
 *
 * Conventions (same as the reference, F:49-61, _core.py:282-288):
 *   - success <=> error_str[0] == '\0'; error_str has room for THINCURR_ERROR_SLEN chars
 *   - strings are NUL-terminated C strings; an empty cache_file means "no cache"
 *   - matrices returned through `void**` are LIBRARY-OWNED host buffers (pinned) in the
 *     reference's column-major layout, valid until the model is destroyed or rebuilt
 *   - LOGICAL(c_bool) arguments are 1-byte bool
 * There is NO CPU fallback: every build routine fails (error_str) if no CUDA device works.
 */
#ifndef THINCURR_B200_H
#define THINCURR_B200_H
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define THINCURR_PATH_SLEN 200  /* OFT_PATH_SLEN  (src/include/local.h) */
#define THINCURR_ERROR_SLEN 200 /* OFT_ERROR_SLEN */

/* ---------------------------------------------------------------------------------------
 * Block 1: reference-compatible entry points (names bound by P:17-120)
 * ------------------------------------------------------------------------------------ */

/* oftpy_init / oftpy_load_xml: src/python/wrappers/oft_base_f.F90:56-110, :117-140.
 * oftpy_load_xml parses <oft><thincurr>...</thincurr></oft> (eta, icoils, vcoils, sens_mask) */
void oftpy_init(int nthreads, bool quiet, const char* input_file, int* slens, void* abort_callback);
void oftpy_load_xml(const char* xml_file, void** oft_node_ptr);
void oftpy_set_nthreads(int nthreads);
void oftpy_set_debug(int debug_level);

/* F:49-228  thincurr_setup: mesh from arrays (np>0; r[np][3], lc[nc][3] 1-based, reg or NULL)
 * or from a native HDF5 mesh file (np<=0).  pmap: periodic map [np] or {-1}.
 * sizes[9] = np,ne,nc,nreg,np_active,nholes,n_vcoils,nelems,n_icoils */
void thincurr_setup(const char* mesh_file, int np, const double* r_loc, int nc, const int* lc_loc,
                    const int* reg_loc, const int* pmap_loc, int jumper_start, void** tw_ptr, int* sizes,
                    char* error_str, void* xml_ptr);

/* F:545-581  self-inductance.  *Lmat_ptr -> double[nelems*nelems], Fortran Lmat(nelems,nelems).
 * use_hodlr=true is rejected (compressed path is out of scope). */
void thincurr_Lmat(void* tw_ptr, bool use_hodlr, void** Lmat_ptr, const char* cache_file, char* error_str);

/* F:585-617  B-field reconstruction operators. *Bmat_ptr -> Bel(nelems,np,3), *Bdr_ptr -> Bdr(np,n_icoils,3) */
void thincurr_Bmat(void* tw_ptr, void* hodlr_ptr, void** Bmat_ptr, void** Bdr_ptr, const char* cache_file,
                   char* error_str);

/* F:621-638  element<->Icoil mutuals. *Mc_ptr -> Ael2dr(nelems,n_icoils).  Also builds Ael2coil/Acoil2coil. */
void thincurr_Mcoil(void* tw_ptr, void** Mc_ptr, const char* cache_file, char* error_str);

/* F:642-690  sensors. *Ms_ptr -> Ael2sen(nsensors,nelems), *Msc_ptr -> Adr2sen(nsensors,n_icoils) */
void thincurr_Msensor(void* tw_ptr, const char* sensor_file, void** Ms_ptr, void** Msc_ptr, int* nsensors,
                      int* njumpers, void** sensor_ptr, const char* cache_file, char* error_str);
/* F:694-703 */
void thincurr_get_sensor_name(void* sensor_ptr, int sensor_ind, char* sensor_name, char* error_str);

/* F:501-521  mutual inductance between two models into CALLER-owned Mmat, Fortran (nelems2,nelems1) */
void thincurr_cross_coupling(void* tw_ptr1, void* tw_ptr2, double* Mmat, const char* cache_file, char* error_str);

/* F:906-921  resistance matrix, 1-based CSR, library-owned (CPU: O(N), thin_wall.F90:1690-1930) */
void thincurr_Rmat(void* tw_ptr, int** kr_ptr, int** lc_ptr, double** mat_ptr, char* error_str);

/* eta accessors used by the reference tests (F:925-1000 region) */
void thincurr_get_eta(void* tw_ptr, double* eta_surf, char* error_str);
void thincurr_set_eta(void* tw_ptr, const double* eta_surf, const double* eta_vol, const double* thickness,
                      char* error_str);

/* The remaining names the reference's Python layer binds at import (P:21-119), so that the unmodified
 * OpenFUSIONToolkit.ThinCurr package loads against this library.  Answered natively: thincurr_scale_va (F:453-466),
 * thincurr_get_eta_vol (F:721-731), thincurr_get_thickness (F:887-902), thincurr_apply_Lmat (F:470-497: dense mat-vec on
 * the device, vals overwritten), thincurr_eigenvalues (F:975-1013: iterative path = Lanczos on the device-resident L,
 * direct=true is refused), thincurr_cross_eval (F:525-541: matrix-free mutual apply, tw_compute_Lmat_MF on the device),
 * thincurr_reduce_model (F:1208-1247: projections on the device, root-level datasets of the reduced-model file).  The others belong to the reference's downstream solvers / plotting and report
 * "not provided" through error_str (those without an error_str print the message and call the abort callback that
 * oftpy_init received, the reference's oft_abort convention). */
void thincurr_setup_io(void* tw_ptr, const char* basepath, bool save_debug, bool legacy_hdf5, char* error_str);
void thincurr_recon_curr(void* tw_ptr, const double* vals, double* curr, int format);
void thincurr_recon_field(void* tw_ptr, const double* pot, const double* coils, double* field, void* hodlr_ptr);
void thincurr_save_field(void* tw_ptr, const double* vals, const char* fieldname);
void thincurr_save_scalar(void* tw_ptr, const double* vals, const char* fieldname);
void thincurr_scale_va(void* tw_ptr, double* vals, bool div_flag);
void thincurr_apply_Lmat(void* tw_ptr, double* vals, void* hodlr_ptr);
void thincurr_cross_eval(void* tw_ptr1, void* tw_ptr2, int nrhs, const double* vec1, double* vec2, char* error_str);
void thincurr_get_eta_vol(void* tw_ptr, double* eta_vol, char* error_str);
void thincurr_get_thickness(void* tw_ptr, double* thickness, char* error_str);
void thincurr_curr_regmat(void* tw_ptr, double* Rmat, char* error_str);
void thincurr_eigenvalues(void* tw_ptr, bool direct, int neigs, double* eig_vals, double* eig_vec, void* hodlr_ptr,
                          char* error_str);
void thincurr_freq_response(void* tw_ptr, bool direct, int fr_limit, double freq, double* fr_driver, void* hodlr_ptr,
                            char* error_str);
void thincurr_time_domain(void* tw_ptr, bool direct, double dt, int nsteps, double cg_atol, double cg_rtol, bool timestep_cn,
                          int nstatus, int nplot, const double* vec_ic, void* sensor_ptr, int ncurr, const double* curr_ptr,
                          int nvolt, const double* volt_ptr, bool volts_full, void* sensor_vals_ptr, void* hodlr_ptr,
                          char* error_str);
void thincurr_time_domain_plot(void* tw_ptr, bool compute_B, bool rebuild_sensors, int nsteps, int nplot, void* sensor_ptr,
                               const double* sensor_vals, int nsensors, void* hodlr_ptr, char* error_str);
void thincurr_reduce_model(void* tw_ptr, const char* filename, int neigs, const double* eig_vec, bool compute_B,
                           void* sensor_ptr, void* hodlr_ptr, char* error_str);

/* Names the reference's BASE package binds at import (src/python/OpenFUSIONToolkit/_interface.py:114-132): mesh
 * objects of the other physics modules.  Exported so that the unmodified package loads; they report "not provided". */
void oft_setup_smesh(int ndim, int np, const double* r_loc, int npc, int nc, const int* lc_loc, const int* reg_loc,
                     int* nregs, void** mesh_ptr);
void oft_smesh_get(void* mesh_ptr, int* ndim, int* np, double** r_loc, int* npc, int* nc, int** lc_loc, int** reg_loc,
                   int* nregs, char* error_str);
void oft_setup_vmesh(int np, const double* r_loc, int npc, int nc, const int* lc_loc, const int* reg_loc, int* nregs,
                     void** mesh_ptr);
void oft_vmesh_get(void* mesh_ptr, int* np, double** r_loc, int* npc, int* nc, int** lc_loc, int** reg_loc, int* nregs,
                   char* error_str);
void dump_cov(void);

/* ---------------------------------------------------------------------------------------
 * Block 2: B200-native flat interface (what a Fortran host binds through ISO_C_BINDING;
 * see include/thincurr_b200_f.F90 and INTEGRATION.md).  All return 0 on success, else an
 * error code with a message retrievable by thincurr_b200_last_error().
 * ------------------------------------------------------------------------------------ */
const char* thincurr_b200_last_error(void);
int thincurr_b200_device_count(void);
void thincurr_b200_destroy(void* tw_ptr);

/* Setup from arrays with explicit node/side sets (what the Python host passes after reading
 * the mesh file itself).  lc 1-based.  nodeset_ptr[nnodesets+1] offsets into nodeset_val (1-based
 * vertex ids); closures = 1-based cell ids (sideset 1) or NULL. */
int thincurr_b200_setup(int np, const double* r, int nc, const int* lc, const int* reg, const int* pmap,
                        int nnodesets, const int* nodeset_ptr, const int* nodeset_val, int nclosures,
                        const int* closures, void* xml_ptr, void** tw_ptr, int* sizes);

/* Model from the arrays a Fortran host already holds in its tw_type (thin_wall.F90:111-154), i.e.
 * what tw_compute_LmatDirect reads: r(3,np), lc(3,nc) AFTER orientation sync (1-based), reg(nc) or
 * NULL, pmap(np) (1-based DOF id, 0 = inactive), np_active, nholes, the hole CSR kfh(nc+1) (1-based
 * Fortran offsets) / lfh(2,nfh) = (signed hole id, 1-based local vertex), and optionally the host's
 * own ca(nc) / qbasis(3,3,nc) (NULL = recomputed with the reference formulas).  No setup work
 * (orientation sync, holes, DOF map) is repeated. */
int thincurr_b200_model_from_tw(int np, const double* r, int nc, const int* lc, const int* reg, const int* pmap,
                                int np_active, int nholes, const int* kfh, const int* lfh, const double* ca,
                                const double* qbasis, void** tw_ptr);

/* Drop-in for tw_compute_LmatDirect(self, Lmat) (thin_wall.F90:887-1186): the full self-inductance
 * matrix into CALLER-owned host memory Lmat(nelems,nelems), rows sharded over all visible devices. */
int thincurr_b200_Lmat_host(void* tw_ptr, double* Lmat);

/* Coil sets / sensors from memory (alternative to XML / floops.loc).  kind: 0 = Vcoil, 1 = Icoil.
 * set_ptr[nsets+1] -> filament ranges; fil_ptr[nfil+1] -> point ranges into pts[][3]. */
int thincurr_b200_set_coils(void* tw_ptr, int kind, int nsets, const int* set_ptr, const int* fil_ptr,
                            const double* pts, const double* scales, const double* radius,
                            const double* res_per_len, const int* sens_mask, int* sizes);
int thincurr_b200_set_sensors(void* tw_ptr, int nsensors, const int* fil_ptr, const double* pts,
                              const double* scale_fac, void** sensor_ptr);

/* Sensor mutuals for sensors given in memory: *Ms_ptr -> Ael2sen(nsensors,nelems), *Msc_ptr -> Adr2sen */
int thincurr_b200_msensor(void* tw_ptr, void* sensor_ptr, void** Ms_ptr, void** Msc_ptr);

/* Row partition of the dense operators.  Rows are grouped in locality-preserving patches
 * (internal order); a shard is a contiguous patch range balanced by pair count.
 * row_ids[nrows] returns the reference (0-based) DOF id of every local row. */
int thincurr_b200_plan(void* tw_ptr, int nshards, int shard, int* nrows);
int thincurr_b200_shard_rows(void* tw_ptr, int nshards, int shard, int* row_ids);
/* info[8]: patch size, patches, chunks, sum of patch cells (halo included), self tiles, chunk pairs,
 * cell pairs of those tiles (diagonal tiles counted in full), vertex patches */
int thincurr_b200_plan_info(void* tw_ptr, int64_t* info);
/* Banded plan of the streamed single-device build behind thincurr_Lmat (large models, one device, page-locked host
 * matrix; replaces the blocking row-by-row fill of thincurr_f.F90:545-581): the vertex DOFs are cut into *nbands ranges of
 * reference ids [band_ref_ptr[b], band_ref_ptr[b+1]) whose patches are [band_patch_ptr[b], band_patch_ptr[b+1]); band b
 * leaves the device as rows [R0,R1) x columns [R0,N) and rows [R1,N) x columns [R0,R1) while the later bands are
 * evaluated.  *nbands = 0: this model is built the ordinary way (small, V-coils, or a reference numbering without
 * locality).  Arrays hold up to 33 entries (may be NULL).  Host-only call (plans, launches nothing); the model keeps
 * the plan. */
int thincurr_b200_stream_plan(void* tw_ptr, int* nbands, int* band_ref_ptr, int* band_patch_ptr);
/* introspection: patch_chunk_ptr[npatch+1]; chunk_info[nchunk][6] = centre xyz, bounding radius, longest edge, cells */
int thincurr_b200_plan_chunks(void* tw_ptr, int* patch_chunk_ptr, double* chunk_info);
/* host->device bytes of one upload of the model (plan mirror) to a device; 0 before the first build */
int64_t thincurr_b200_model_bytes(void* tw_ptr);

/* Self-inductance rows of one shard into caller-provided DEVICE memory d_out[nrows][ld]
 * (row r = full reference row row_ids[r], i.e. Lmat(:,row_ids[r]+1); ld >= nelems), on the
 * CUDA device that owns d_out, enqueued on `stream` (cudaStream_t as void*), asynchronous.
 * stats[8] (host, optional; forces a stream sync): [0] far pairs evaluated, [1] near T evaluations,
 * [2] 1/r evaluations, [3] analytic-potential evaluations; the _host variant adds [5] host->device
 * bytes of the model upload and [6] device->host bytes of the rows. */
int thincurr_b200_Lmat_shard(void* tw_ptr, int nshards, int shard, double* d_out, int64_t ld, void* stream,
                             int64_t* stats);
/* Symmetric multi-device build: every pair integral is evaluated on exactly one device.  The row partition gives every
 * shard the same work; shard s builds d_out[nrows][ld] for its rows: its diagonal block completely, and of every block it
 * shares with another shard the tiles (row patch pa, column patch pb) with ((pa + pb) even) == (pa < pb) -- a checkerboard,
 * half of the block; the other entries are left zero and are the transposes of entries the other shard computed (L is
 * symmetric, thin_wall.F90:1146-1151).  thincurr_b200_Lmat_exchange (block 3) completes them after the assembly by
 * reading the peers' rows.  thincurr_b200_shard_rows_sym: row count and/or reference DOF ids of the rows of a shard (either
 * pointer may be NULL); thincurr_b200_dof_patches: patch index of every vertex / hole DOF (reference id). */
int thincurr_b200_shard_rows_sym(void* tw_ptr, int nshards, int shard, int* nrows, int* row_ids);
int thincurr_b200_dof_patches(void* tw_ptr, int nshards, int* patch_of_dof);
int thincurr_b200_Lmat_shard_sym(void* tw_ptr, int nshards, int shard, double* d_out, int64_t ld, void* stream,
                                 int64_t* stats);
/* Same from HOST mesh each call (uploads model, builds, copies rows back to h_out[nrows][ld]);
 * the end-to-end path used by bench.py's e2e leg. */
int thincurr_b200_Lmat_shard_host(void* tw_ptr, int nshards, int shard, double* h_out, int64_t ld,
                                  int64_t* stats);
/* Dense block of the self-inductance matrix: d_out[i][j] = Lmat(col_ids[j]+1, row_ids[i]+1) (= L(row_ids[i], col_ids[j]),
 * the matrix is symmetric) for arbitrary subsets of the vertex/hole DOFs (0-based reference ids, no duplicates), into
 * caller-provided DEVICE memory d_out[nrows][ld >= ncols], asynchronous on `stream`.  Same pair integrals and role rule
 * as the full build; this is the evaluator a hierarchical (HODLR) compression calls for its dense diagonal and
 * near-field blocks (tw_compute_Lmatblock, thin_wall_hodlr.F90:136-404).  Work is done per (row patch, column patch)
 * tile, so thin strips cost as much as the patches they touch. */
int thincurr_b200_Lmat_block(void* tw_ptr, int nrows, const int* row_ids, int ncols, const int* col_ids, double* d_out,
                             int64_t ld, void* stream);
/* HODLR dense-block builders with the reference's own semantics.  Blocks are VERTEX subsets (0-based mesh
 * vertex ids, no duplicates: oft_tw_block%ipts); the cells touching them (icell) and the inverse map are derived here.
 * `out` may be HOST or DEVICE memory (detected); host output makes the call synchronous.  The device-side cell arrays of
 * a model are uploaded on the first call and kept (ACA+ asks for thousands of one-row strips, thin_wall_hodlr.F90:1224-1428).
 *  - Lmatblock = tw_compute_Lmatblock (thin_wall_hodlr.F90:289-404): out[a][b] = Lmat(col_pts[b], row_pts[a]); the ROW
 *    block's cell is the analytic side of every near pair, no pair is skipped, vertex DOFs only.  tw_col NULL = tw_row.
 *  - LmatHole = tw_compute_LmatHole(self,self,...) (:136-285): out[h][:] = Lmat(:, h), h over the nholes + n_vcoils
 *    hole / V-coil columns of hole_Vcoil_mat, ld >= nelems.
 *  - Bops_block = tw_compute_Bops_block (:580-691): out[a][b] = Bop(col_pts[b], row_pts[a]) for component dir = 0,1,2;
 *    dir < 0: all three in one sweep, out[3][nrp][ld].
 * Work per call: (cells of the row block) x (cells or vertices of the column block) pair integrals, nothing else. */
int thincurr_b200_Lmatblock(void* tw_row, void* tw_col, int nrp, const int* row_pts, int ncp, const int* col_pts, double* out,
                            int64_t ld, void* stream);
int thincurr_b200_LmatHole(void* tw_ptr, double* out, int64_t ld, void* stream);
int thincurr_b200_Bops_block(void* tw_ptr, int nrp, const int* row_pts, int ncp, const int* col_pts, int dir, double* out,
                             int64_t ld, void* stream);
/* Matrix-free apply between two models = tw_compute_Lmat_MF (thin_wall.F90:1190-1414; what thincurr_cross_eval calls):
 * vec2[q][:] = M vec1[q][:] for host arrays vec1[nrhs][nelems1], vec2[nrhs][nelems2]; counts[3] (optional) returns the
 * number of cell pairs per class of its quadrature heuristic (far / close / very close). */
int thincurr_b200_cross_eval(void* tw_ptr1, void* tw_ptr2, int nrhs, const double* vec1, double* vec2, int64_t* counts);
/* Minimal HDF5 writer (no libhdf5 needed): root-level contiguous little-endian datasets, float64 (is_f64[i] != 0) or
 * int32, dims[] = the dimensions of all items concatenated in C order (slowest first); at most 32 items (one symbol-table
 * node: group leaf K = 4 up to 8 items, the libhdf5 default, 16 above).  This is the
 * container of the reference's Bmat cache (thin_wall.F90:2208-2225: MODEL_hash, Bel_X|Y|Z, Bdr_X|Y|Z), which
 * thincurr_Bmat writes and reads through it. */
int thincurr_b200_h5_write(const char* path, int nitems, const char* const* names, const int* is_f64, const int* ranks,
                           const int64_t* dims, const void* const* data);
/* B-field operator rows (element index sharded the same way): d_out[3][np][nrows] */
int thincurr_b200_Bel_shard(void* tw_ptr, int nshards, int shard, double* d_out, void* stream);

/* iquad histogram / visited-pair count of the reference loop computed on the GPU:
 * hist[19], visited = # ordered pairs not skipped by thin_wall.F90:1034. */
int thincurr_b200_pair_stats(void* tw_ptr, int64_t* hist, int64_t* visited);

/* Number of operator kernels this library has launched so far (all devices, this process). */
long long thincurr_b200_launch_count(void);

/* FP64 DFMA peak microbenchmark on the current device (TFLOP/s, FMA = 2 flops). */
double thincurr_b200_dfma_peak(int device, double* sm_clock_mhz);

/* Introspection for tests. */
int thincurr_b200_get_model(void* tw_ptr, int* pmap, int* lc, int* kfh, int* lfh, double* qbasis, double* ca);
int thincurr_b200_hashes(void* tw_ptr, int32_t* hash_lc, int32_t* hash_r);

/* Release the device-side state of a model (plan mirrors, row-block scratch of the host-buffer entry points). */
int thincurr_b200_release_device(void* tw_ptr);

/* ---------------------------------------------------------------------------------------
 * Block 3: multi-device data plane of the sharded operators (csrc/tw_shard.cu, csrc/tw_solve.cu).
 * The assembly needs no inter-device traffic; these entry points are what follows it.
 * ------------------------------------------------------------------------------------ */
/* Device memory that other ranks of the same node can map (cudaMalloc + cudaIpc*), and peer access between the
 * devices of one process. handle64 = 64 bytes (cudaIpcMemHandle_t). */
int thincurr_b200_device_alloc(int64_t bytes, void** d_ptr);
int thincurr_b200_device_free(void* d_ptr);
int thincurr_b200_ipc_export(void* d_ptr, unsigned char* handle64);
int thincurr_b200_ipc_open(const unsigned char* handle64, void** d_ptr);
int thincurr_b200_ipc_close(void* d_ptr);
int thincurr_b200_enable_peer(int peer_device);

/* Exchange after a symmetric build (thin_wall.F90:1146-1151 across shards): shard `shard` fills the entries of its rows
 * d_out[nrows][ld] that the other shards evaluated (see thincurr_b200_Lmat_shard_sym) with their transposes, READING
 * peer_rows[s] (device pointer to shard s's row block with the same ld: a peer device of this process or a cudaIpc-mapped
 * pointer of another rank; needed for every s != shard) over NVLink.  Asynchronous on `stream`; the caller orders it after
 * the peers' builds (event / stream-ordered collective).  Peers may run their own exchange at the same time (the entries
 * read here are not the ones they write). */
int thincurr_b200_Lmat_exchange(void* tw_ptr, int nshards, int shard, double* d_out, int64_t ld,
                                const double* const* peer_rows, void* stream);
/* "One gather over NVLink when the full matrix is requested on one device": rows of all shards (pointers readable from
 * the current device, ld_src) into d_full[nelems][ld_full] in the reference row order.  sym != 0: the shards are the
 * symmetric partition's (already exchanged). */
int thincurr_b200_Lmat_gather(void* tw_ptr, int nshards, int sym, const double* const* shard_rows, int64_t ld_src,
                              double* d_full, int64_t ld_full, void* stream);
/* Export of a shard's rows (device) into a host matrix h_full[nelems][ld_full] in the reference layout, streamed
 * through pinned buffers (matrices that fit no single device: 150k-vertex vessel, 180 GB).  h_full == NULL: stream only
 * (the rows pass through the pinned staging buffers and are dropped; measures what an export sees). */
int thincurr_b200_rows_to_host(void* tw_ptr, int nshards, int shard, int sym, const double* d_rows, int64_t ld,
                               double* h_full, int64_t ld_full);
/* Export into the reference's `Lmat.save` cache (thin_wall.F90:1161-1171): _begin writes the header record and sizes the
 * file, _rows writes the upper-packed records of a shard's rows at their file offsets (ranks write concurrently). */
int thincurr_b200_Lmat_save_begin(void* tw_ptr, const char* path);
int thincurr_b200_Lmat_save_rows(void* tw_ptr, const char* path, int nshards, int shard, int sym, const double* d_rows,
                                 int64_t ld);

/* Dense apply on a resident row block: d_y[nrows] = d_rows[nrows][ld] . d_x[n] (thincurr_apply_Lmat, F:470-497, per
 * shard; the caller all-gathers y).  HBM-bound row kernel, asynchronous on `stream`. */
int thincurr_b200_rows_apply(const double* d_rows, int64_t ld, int nrows, int n, const double* d_x, double* d_y,
                             void* stream);
/* Leading `neigs` eigenvalues of L x = lambda R x (lr_eigenmodes_arpack, thin_wall_solvers.F90:119-224) by Lanczos in
 * the R inner product; y = L x is supplied by the caller (HOST vectors of nelems doubles; returns 0 on success) so the
 * matrix can stay sharded over devices / ranks.  eig_vec[neigs][nelems]; tol = relative residual (<=0: 1e-10);
 * max_dim = basis limit (<=0: 400); n_applies (optional) returns the number of mat-vecs used. */
typedef int (*thincurr_b200_apply_fn)(void* user, const double* x, double* y);
int thincurr_b200_lr_eigs(void* tw_ptr, int neigs, double tol, int max_dim, thincurr_b200_apply_fn apply, void* user,
                          double* eig_vals, double* eig_vec, int* n_applies);

#ifdef __cplusplus
}
#endif
#endif
