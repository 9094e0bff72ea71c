/* thincurr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP) of the ThinCurr dense operator builds of the
 * reference (OpenFUSIONToolkit @ d08f001b, src/physics/thin_wall.F90).  It exists only
 * to check the CUDA path (tests/, __graft_entry__.smoke()) and to be timed as the
 * "reference-equivalent CPU build" (bench.py cpu_baseline / --impl reference).  Nothing
 * under openfusiontoolkit_b200/ may link, import or call it.
 *
 * Parity status: the reference itself cannot be built here (no Fortran compiler, no
 * HDF5), so this port is pinned against the reference's own regression goldens
 * (src/tests/physics/test_ThinCurr.py: plate/cyl/torus/passive eigenvalues :984,:1029,
 * :1076,:1164 and the plate frequency response :1004-1005) and the quadrature KAT
 * (src/tests/grid/quad_2d.tests) -- see tests/test_oracle_golden.py.  Entry-wise values
 * of L/Bel/M are NOT pinned by any reference test.
 *
 * Loop structure, operation order and OpenMP strategy follow the reference line by
 * line so that it is a fair CPU baseline: `schedule(dynamic,100)` over row cells with
 * atomic scatter (thin_wall.F90:1008-1126).  Build with the reference's release flags
 * (-O2 -fopenmp, no fast-math, no -march=native; src/CMakeLists.txt:182,231-238).
 *
 * All index arrays are 0-based here except `pmap` (1-based DOF id, 0 = inactive) and
 * the hole id in `lfh` (signed, 1-based), mirroring the Fortran semantics.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "quad_tables.h"

#define TARGET_ERR 1.0e-8 /* thin_wall.F90:158 */
static const double PI = 3.14159265358979323846;

typedef struct {
  int np, nc, np_active, nholes, n_vcoils, n_icoils, nelems, nfh;
  const double *r;      /* [np][3]                       mesh%r        */
  const int *lc;        /* [nc][3] 0-based               mesh%lc       */
  const int *reg;       /* [nc] 1-based                  mesh%reg      */
  const double *ca;     /* [nc]                          mesh%ca       */
  const double *va;     /* [np]                          mesh%va       */
  const double *norm;   /* [nc][3] unit normals          trimesh_norm  */
  const double *qbasis; /* [nc][3 vert][3 xyz]           tw%qbasis     */
  const int *pmap;      /* [np] 1-based DOF, 0 inactive  tw%pmap       */
  const int *kfh;       /* [nc+1] 0-based offsets        tw%kfh        */
  const int *lfh;       /* [nfh][2] (+-hole id, local v) tw%lfh        */
  const int *sens_mask; /* [nreg] or NULL                tw%sens_mask  */
} tco_model;

/* A list of coil sets; each set is a list of filaments (polylines). */
typedef struct {
  int nsets;
  const int *set_ptr;     /* [nsets+1] filament range of each set */
  const int *fil_ptr;     /* [nfil+1] point range of each filament */
  const double *pts;      /* [npts_total][3] */
  const double *scales;   /* [nfil] */
  const double *radius;   /* [nfil] */
  const int *sens_mask;   /* [nsets] */
} tco_coils;

static inline int isign(int v) { return v < 0 ? -1 : 1; }

/* ------------------------------------------------------------------ */
/* tw_compute_phipot, thin_wall.F90:1934-1985                          */
/* ------------------------------------------------------------------ */
static inline void cross3(const double *a, const double *b, double *c) {
  /* oft_local.F90:314-320 */
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double *a, const double *b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

double tco_phipot(const double pt_cell[3][3], const double pt[3]) {
  double e1[3], e2[3], nhat[3], r[3][3], rmag[3], c[3][3], gam[3], tmpv[3];
  for (int d = 0; d < 3; d++) {
    e1[d] = pt_cell[1][d] - pt_cell[0][d];
    e2[d] = pt_cell[2][d] - pt_cell[1][d];
  }
  cross3(e1, e2, nhat);
  double nmag = sqrt(dot3(nhat, nhat));
  for (int d = 0; d < 3; d++) nhat[d] = nhat[d] / nmag;
  for (int i = 0; i < 3; i++) {
    for (int d = 0; d < 3; d++) r[i][d] = pt_cell[i][d] - pt[d];
    rmag[i] = sqrt(dot3(r[i], r[i]));
    double dn = dot3(nhat, r[i]);
    for (int d = 0; d < 3; d++) c[i][d] = r[i][d] - dn * nhat[d];
  }
  cross3(r[1], r[2], tmpv);
  double num = dot3(r[0], tmpv);
  double den = rmag[0] * rmag[1] * rmag[2] + dot3(r[0], r[1]) * rmag[2] +
               dot3(r[0], r[2]) * rmag[1] + dot3(r[1], r[2]) * rmag[0];
  double omega = 2.0 * atan2(num, den);
  for (int i = 0; i < 3; i++) {
    const double *yp = r[i];
    const double *yp1 = r[(i + 1) % 3];
    double dv[3] = {yp1[0] - yp[0], yp1[1] - yp[1], yp1[2] - yp[2]};
    double tmp = sqrt(dot3(dv, dv));
    double n2 = sqrt(dot3(yp1, yp1)) * tmp + dot3(yp1, dv);
    double d2 = sqrt(dot3(yp, yp)) * tmp + dot3(yp, dv);
    if (fabs(d2) < 1.e-14 || tmp < 1.e-14)
      gam[i] = 0.0;
    else
      gam[i] = log(n2 / d2) / tmp;
  }
  double phi = 0.0;
  for (int i = 0; i < 3; i++) {
    cross3(c[i], c[(i + 1) % 3], tmpv);
    phi = phi + dot3(nhat, tmpv) * gam[i];
  }
  phi = phi - dot3(nhat, r[0]) * omega;
  return phi;
}

/* Order selection, thin_wall.F90:1055-1059 (same expression at :677-681,
 * :1535-1539, :2041-2047).  Fortran INT() truncates toward zero; ln(0)=-inf gives
 * 0 -> clamped to 4.  The clamp is applied in double to avoid int overflow UB for
 * ratio < ~1e-8 (unreachable for sane meshes). */
static inline int tco_iquad(double dl_min, double dl_max) {
  if (dl_min < 1.e-8) return 18;
  double q = fabs(trunc(log(TARGET_ERR) / log(1.0 - dl_min / dl_max)));
  if (!(q < 18.0)) return 18;
  int iq = (int)q;
  return iq < 4 ? 4 : iq;
}

static inline void quad_point(int iquad, int q, const double P[3][3], double *x) {
  const double *b = TCQ_PTS[TCQ_OFF[iquad] + q];
  for (int d = 0; d < 3; d++) x[d] = b[0] * P[0][d] + b[1] * P[1][d] + b[2] * P[2][d];
}

/* Pair integral T(i,j), thin_wall.F90:1044-1083.  i = analytic side when near. */
double tco_pair_T(const double pts_i[3][3], double area_i, const double pts_j[3][3],
                  double area_j, int *iquad_out) {
  double dl_min = 1.e99;
  double dl_max = sqrt(fmax(area_i, area_j) * 2.0);
#ifdef TCO_SIMD /* CPU-baseline build only: the reference's `!$omp simd ... reduction` (thin_wall.F90:1047) */
#pragma omp simd reduction(max : dl_max) reduction(min : dl_min)
#endif
  for (int ii = 0; ii < 3; ii++)
    for (int jj = 0; jj < 3; jj++) {
      double dx = pts_i[ii][0] - pts_j[jj][0], dy = pts_i[ii][1] - pts_j[jj][1],
             dz = pts_i[ii][2] - pts_j[jj][2];
      double d = sqrt(dx * dx + dy * dy + dz * dz);
      dl_min = fmin(dl_min, d);
      dl_max = fmax(dl_max, d);
    }
  int iquad = tco_iquad(dl_min, dl_max);
  if (iquad_out) *iquad_out = iquad;
  int nq = TCQ_NP[iquad];
  const double *w = TCQ_WTS + TCQ_OFF[iquad];
  double tmp = 0.0;
  if (iquad > 10) {
    for (int jj = 0; jj < nq; jj++) {
      double pt_j[3];
      quad_point(iquad, jj, pts_j, pt_j);
      tmp = tmp + w[jj] * tco_phipot(pts_i, pt_j);
    }
    tmp = tmp * area_j;
  } else {
#ifdef TCO_SIMD /* CPU-baseline build only: `!$omp simd collapse(1) private(pt_i,pt_j) reduction(+:tmp)` (thin_wall.F90:1070);
                   the parity oracle keeps the sequential summation order */
#pragma omp simd reduction(+ : tmp)
#endif
    for (int ii = 0; ii < nq; ii++) {
      double pt_i[3];
      quad_point(iquad, ii, pts_i, pt_i);
      for (int jj = 0; jj < nq; jj++) {
        double pt_j[3];
        quad_point(iquad, jj, pts_j, pt_j);
        double dx = pt_i[0] - pt_j[0], dy = pt_i[1] - pt_j[1], dz = pt_i[2] - pt_j[2];
        tmp = tmp + w[jj] * w[ii] / sqrt(dx * dx + dy * dy + dz * dz);
      }
    }
    tmp = tmp * area_i * area_j;
  }
  return tmp;
}

static inline void load_cell(const tco_model *m, int c, double P[3][3], double E[3][3]) {
  for (int k = 0; k < 3; k++) {
    int v = m->lc[3 * c + k];
    for (int d = 0; d < 3; d++) {
      P[k][d] = m->r[3 * v + d];
      E[k][d] = m->qbasis[9 * c + 3 * k + d];
    }
  }
}

/* ------------------------------------------------------------------ */
/* tw_compute_LmatDirect, thin_wall.F90:887-1186                        */
/*   Lmat is Fortran (col%nelems,row%nelems): Lmat[ik*ncol + jk].       */
/*   col == NULL -> self inductance (needs Ael2coil/Acoil2coil if       */
/*   n_vcoils>0).  [i_begin,i_end) restricts the outer row-cell loop    */
/*   (bench sampling); finalize!=0 runs the V-coil fill, mirror and     */
/*   1/(4 pi) scaling (:1128-1154).  hist[19] (optional) = iquad        */
/*   histogram of visited pairs.  Returns # visited pairs.              */
/* ------------------------------------------------------------------ */
long long tco_lmat_direct(const tco_model *row, const tco_model *col, double *Lmat,
                          const double *Ael2coil, const double *Acoil2coil, int i_begin,
                          int i_end, int finalize, long long *hist) {
  const int Lself = (col == NULL);
  if (Lself) col = row;
  const long long ncol = col->nelems;
  long long visited = 0;
  long long hloc[19];
  memset(hloc, 0, sizeof(hloc));
#pragma omp parallel reduction(+ : visited) reduction(+ : hloc[:19])
  {
#pragma omp for schedule(dynamic, 100)
    for (int i = i_begin; i < i_end; i++) {
      double pts_i[3][3], evec_i[3][3], pts_j[3][3], evec_j[3][3];
      double area_i = row->ca[i];
      load_cell(row, i, pts_i, evec_i);
      int imin = 0;
      if (Lself) {
        imin = row->pmap[row->lc[3 * i]];
        for (int k = 1; k < 3; k++) {
          int p = row->pmap[row->lc[3 * i + k]];
          if (p < imin) imin = p;
        }
        for (int ii = row->kfh[i]; ii < row->kfh[i + 1]; ii++) {
          int h = abs(row->lfh[2 * ii]) + row->np_active;
          if (h < imin) imin = h;
        }
      }
      for (int j = 0; j < col->nc; j++) {
        if (Lself) {
          int jmax = col->pmap[col->lc[3 * j]];
          for (int k = 1; k < 3; k++) {
            int p = col->pmap[col->lc[3 * j + k]];
            if (p > jmax) jmax = p;
          }
          for (int jj = col->kfh[j]; jj < col->kfh[j + 1]; jj++) {
            int h = abs(col->lfh[2 * jj]) + col->np_active;
            if (h > jmax) jmax = h;
          }
          if (jmax < imin) continue;
        }
        double area_j = col->ca[j];
        load_cell(col, j, pts_j, evec_j);
        int iquad;
        double tmp = tco_pair_T(pts_i, area_i, pts_j, area_j, &iquad);
        visited++;
        hloc[iquad]++;
        /* scatter, :1085-1124 */
        for (int ii = 0; ii < 3; ii++) {
          int ik = row->pmap[row->lc[3 * i + ii]];
          if (ik == 0) continue;
          for (int jj = 0; jj < 3; jj++) {
            int jk = col->pmap[col->lc[3 * j + jj]];
            if ((Lself && jk < ik) || jk <= 0) continue;
            double v = dot3(evec_i[ii], evec_j[jj]) * tmp;
#pragma omp atomic
            Lmat[(long long)(ik - 1) * ncol + (jk - 1)] += v;
          }
          for (int jj = col->kfh[j]; jj < col->kfh[j + 1]; jj++) {
            int jk = abs(col->lfh[2 * jj]) + col->np_active;
            double v = isign(col->lfh[2 * jj]) * dot3(evec_i[ii], evec_j[col->lfh[2 * jj + 1]]) * tmp;
#pragma omp atomic
            Lmat[(long long)(ik - 1) * ncol + (jk - 1)] += v;
          }
        }
        for (int ii = row->kfh[i]; ii < row->kfh[i + 1]; ii++) {
          int ik = abs(row->lfh[2 * ii]) + row->np_active;
          const double *ei = evec_i[row->lfh[2 * ii + 1]];
          int si = isign(row->lfh[2 * ii]);
          for (int jj = col->kfh[j]; jj < col->kfh[j + 1]; jj++) {
            int jk = abs(col->lfh[2 * jj]) + col->np_active;
            if (Lself && jk < ik) continue;
            double v = si * isign(col->lfh[2 * jj]) * dot3(ei, evec_j[col->lfh[2 * jj + 1]]) * tmp;
#pragma omp atomic
            Lmat[(long long)(ik - 1) * ncol + (jk - 1)] += v;
          }
          if (!Lself) {
            for (int jj = 0; jj < 3; jj++) {
              int jk = col->pmap[col->lc[3 * j + jj]];
              if (jk <= 0) continue;
              double v = si * dot3(ei, evec_j[jj]) * tmp;
#pragma omp atomic
              Lmat[(long long)(ik - 1) * ncol + (jk - 1)] += v;
            }
          }
        }
      }
    }
  }
  if (hist)
    for (int k = 0; k < 19; k++) hist[k] += hloc[k];
  if (finalize) {
    const long long nrow = row->nelems;
    if (Lself) {
      const int ns = row->np_active + row->nholes;
      for (int i = 0; i < ns; i++)
        for (int j = 0; j < row->n_vcoils; j++)
          Lmat[(long long)i * ncol + (ns + j)] = Ael2coil[(long long)j * row->nelems + i]; /* Ael2coil(i,j) */
      for (int i = 0; i < row->n_vcoils; i++)
        for (int j = i; j < row->n_vcoils; j++)
          Lmat[(long long)(ns + i) * ncol + (ns + j)] = Acoil2coil[(long long)j * row->n_vcoils + i]; /* Acoil2coil(i,j) */
      for (long long i = 0; i < nrow; i++)
        for (long long j = 0; j < i; j++) Lmat[i * ncol + j] = Lmat[j * ncol + i];
    }
    const double s = 4.0 * PI;
#pragma omp parallel for
    for (long long k = 0; k < nrow * ncol; k++) Lmat[k] = Lmat[k] / s;
  }
  return visited;
}

/* ------------------------------------------------------------------ */
/* element <-> polyline potential, shared by :660-707 and :1524-1561    */
/* ------------------------------------------------------------------ */
static inline double cell_point_pot(const double pts_i[3][3], double area_i, const double *cpt,
                                    double *dl_min_out) {
  double dl_min = 1.e99;
  double dl_max = sqrt(area_i * 2.0);
  for (int ii = 0; ii < 3; ii++) {
    double dx = pts_i[ii][0] - cpt[0], dy = pts_i[ii][1] - cpt[1], dz = pts_i[ii][2] - cpt[2];
    double d = sqrt(dx * dx + dy * dy + dz * dz);
    dl_min = fmin(dl_min, d);
    dl_max = fmax(dl_max, d);
  }
  if (dl_min_out) *dl_min_out = dl_min;
  int iquad = tco_iquad(dl_min, dl_max);
  if (iquad > 10) return tco_phipot(pts_i, cpt);
  int nq = TCQ_NP[iquad];
  const double *w = TCQ_WTS + TCQ_OFF[iquad];
  double pot = 0.0;
  for (int ii = 0; ii < nq; ii++) {
    double pt_i[3];
    quad_point(iquad, ii, pts_i, pt_i);
    double dx = pt_i[0] - cpt[0], dy = pt_i[1] - cpt[1], dz = pt_i[2] - cpt[2];
    pot = pot + w[ii] / sqrt(dx * dx + dy * dy + dz * dz);
  }
  return pot * area_i;
}

/* tw_compute_Ael2dr hot loop, thin_wall.F90:644-726.
 * out: Ael2coil_tmp Fortran (nelems, ncoils_tot) -> out[j*nelems + (ik-1)]
 * nrad_cross[nsets] counts points closer than the filament radius (:673-676). */
void tco_ael2coil(const tco_model *m, const tco_coils *cs, double *out, int *nrad_cross) {
  const int nsets = cs->nsets;
#pragma omp parallel
  {
    double *atmp = (double *)malloc(sizeof(double) * 3 * (nsets > 0 ? nsets : 1));
#pragma omp for
    for (int i = 0; i < m->nc; i++) {
      double pts_i[3][3], evec_i[3][3];
      double area_i = m->ca[i];
      load_cell(m, i, pts_i, evec_i);
      for (int k = 0; k < 3 * nsets; k++) atmp[k] = 0.0;
      for (int j = 0; j < nsets; j++) {
        for (int k = cs->set_ptr[j]; k < cs->set_ptr[j + 1]; k++) {
          double tmp[3] = {0.0, 0.0, 0.0};
          double pot_last = 0.0;
          int p0 = cs->fil_ptr[k], p1 = cs->fil_ptr[k + 1];
          for (int kk = p0; kk < p1; kk++) {
            const double *cpt = cs->pts + 3 * kk;
            double dl_min;
            double pot_tmp = cell_point_pot(pts_i, area_i, cpt, &dl_min);
            if (dl_min < cs->radius[k]) {
#pragma omp atomic
              nrad_cross[j]++;
            }
            if (kk > p0) {
              double cvec[3] = {cpt[0] - cpt[-3], cpt[1] - cpt[-2], cpt[2] - cpt[-1]};
              for (int jj = 0; jj < 3; jj++)
                tmp[jj] = tmp[jj] + dot3(evec_i[jj], cvec) * (pot_tmp + pot_last) / 2.0;
            }
            pot_last = pot_tmp;
          }
          for (int jj = 0; jj < 3; jj++) atmp[3 * j + jj] = atmp[3 * j + jj] + cs->scales[k] * tmp[jj];
        }
      }
      for (int ii = 0; ii < 3; ii++) {
        int ik = m->pmap[m->lc[3 * i + ii]];
        if (ik == 0) continue;
        for (int j = 0; j < nsets; j++) {
#pragma omp atomic
          out[(long long)j * m->nelems + (ik - 1)] += atmp[3 * j + ii];
        }
      }
      for (int ii = m->kfh[i]; ii < m->kfh[i + 1]; ii++) {
        int ik = abs(m->lfh[2 * ii]) + m->np_active;
        for (int j = 0; j < nsets; j++) {
          double v = isign(m->lfh[2 * ii]) * atmp[3 * j + m->lfh[2 * ii + 1]];
#pragma omp atomic
          out[(long long)j * m->nelems + (ik - 1)] += v;
        }
      }
    }
    free(atmp);
  }
}

/* Filament <-> filament Neumann sums.
 * tw_compute_Lmat_coils :814-849 (regularised, delta = r_c^2/sqrt(e) of the ROW filament)
 * and the coil->sensor loop :1597-1632 (delta = 0, masked sets skipped).
 * rows/cols are coil-set lists; out is Fortran (nrows, ncols): out[j*nrows + l].
 * regularize != 0 -> Lmat_coils flavour (row scale applied, :838); else sensor flavour. */
void tco_filament_mutual(const tco_coils *rows, const tco_coils *cols, int regularize, double *out) {
  const double sqrt_e = sqrt(exp(1.0));
#pragma omp parallel for
  for (int l = 0; l < rows->nsets; l++) {
    double *atmp = (double *)calloc(cols->nsets > 0 ? cols->nsets : 1, sizeof(double));
    for (int i = rows->set_ptr[l]; i < rows->set_ptr[l + 1]; i++) {
      double thick = regularize ? (rows->radius[i] * rows->radius[i]) / sqrt_e : 0.0;
      int p0 = rows->fil_ptr[i], p1 = rows->fil_ptr[i + 1];
      for (int ii = p0 + 1; ii < p1; ii++) {
        const double *pt_i = rows->pts + 3 * ii;
        const double *pt_i_last = rows->pts + 3 * (ii - 1);
        double rvec_i[3] = {pt_i[0] - pt_i_last[0], pt_i[1] - pt_i_last[1], pt_i[2] - pt_i_last[2]};
        for (int j = 0; j < cols->nsets; j++) {
          if (!regularize && cols->sens_mask && cols->sens_mask[j]) continue;
          for (int k = cols->set_ptr[j]; k < cols->set_ptr[j + 1]; k++) {
            double pot_last = 0.0, tmp = 0.0;
            int q0 = cols->fil_ptr[k], q1 = cols->fil_ptr[k + 1];
            for (int kk = q0; kk < q1; kk++) {
              const double *cpt = cols->pts + 3 * kk;
              double a0 = pt_i[0] - cpt[0], a1 = pt_i[1] - cpt[1], a2 = pt_i[2] - cpt[2];
              double b0 = pt_i_last[0] - cpt[0], b1 = pt_i_last[1] - cpt[1], b2 = pt_i_last[2] - cpt[2];
              double pot_tmp = (1.0 / sqrt((a0 * a0 + a1 * a1 + a2 * a2) + thick) +
                                1.0 / sqrt((b0 * b0 + b1 * b1 + b2 * b2) + thick)) / 2.0;
              if (kk > q0) {
                double cvec[3] = {cpt[0] - cpt[-3], cpt[1] - cpt[-2], cpt[2] - cpt[-1]};
                tmp = tmp + dot3(rvec_i, cvec) * (pot_tmp + pot_last) / 2.0;
              }
              pot_last = pot_tmp;
            }
            if (regularize)
              atmp[j] = atmp[j] + cols->scales[k] * rows->scales[i] * tmp;
            else
              atmp[j] = atmp[j] + cols->scales[k] * tmp;
          }
        }
      }
    }
    for (int j = 0; j < cols->nsets; j++) out[(long long)j * rows->nsets + l] = atmp[j];
    free(atmp);
  }
}

/* tw_compute_mutuals element->sensor loop, thin_wall.F90:1508-1583.
 * sensors are single-filament "sets" (scales[] = scale_fac).
 * out: Ael2sen Fortran (nsensors, nelems) -> out[(ik-1)*nsens + j]; scale_fac applied (:1581). */
void tco_ael2sen(const tco_model *m, const tco_coils *sens, double *out) {
  const int ns = sens->nsets;
#pragma omp parallel
  {
    double *atmp = (double *)malloc(sizeof(double) * 3 * (ns > 0 ? ns : 1));
#pragma omp for
    for (int i = 0; i < m->nc; i++) {
      if (m->sens_mask && m->sens_mask[m->reg[i] - 1]) continue;
      double pts_i[3][3], evec_i[3][3];
      double area_i = m->ca[i];
      load_cell(m, i, pts_i, evec_i);
      for (int k = 0; k < 3 * ns; k++) atmp[k] = 0.0;
      for (int j = 0; j < ns; j++) {
        double pot_last = 0.0;
        int f = sens->set_ptr[j];
        int p0 = sens->fil_ptr[f], p1 = sens->fil_ptr[f + 1];
        for (int jj = p0; jj < p1; jj++) {
          const double *pt_j = sens->pts + 3 * jj;
          double tmp = cell_point_pot(pts_i, area_i, pt_j, NULL);
          if (jj > p0) {
            double rvec_j[3] = {pt_j[0] - pt_j[-3], pt_j[1] - pt_j[-2], pt_j[2] - pt_j[-1]};
            for (int ik = 0; ik < 3; ik++)
              atmp[3 * j + ik] = atmp[3 * j + ik] + dot3(rvec_j, evec_i[ik]) * (tmp + pot_last) / 2.0;
          }
          pot_last = tmp;
        }
      }
      for (int ii = 0; ii < 3; ii++) {
        int ik = m->pmap[m->lc[3 * i + ii]];
        if (ik == 0) continue;
        for (int j = 0; j < ns; j++) {
#pragma omp atomic
          out[(long long)(ik - 1) * ns + j] += atmp[3 * j + ii];
        }
      }
      for (int ii = m->kfh[i]; ii < m->kfh[i + 1]; ii++) {
        int ik = abs(m->lfh[2 * ii]) + m->np_active;
        for (int j = 0; j < ns; j++) {
          double v = isign(m->lfh[2 * ii]) * atmp[3 * j + m->lfh[2 * ii + 1]];
#pragma omp atomic
          out[(long long)(ik - 1) * ns + j] += v;
        }
      }
    }
    free(atmp);
  }
  for (int j = 0; j < ns; j++) {
    double sf = sens->scales[sens->set_ptr[j]];
    for (long long e = 0; e < m->nelems; e++) out[e * ns + j] = out[e * ns + j] * sf;
  }
}

/* ------------------------------------------------------------------ */
/* tw_compute_Bops element part, thin_wall.F90:2014-2112 (no 1/4pi).    */
/*   Bel Fortran (nelems, np, 3): Bel[(jj*np + j)*nelems + (ik-1)].     */
/*   [i_begin,i_end) restricts the cell loop (bench sampling).          */
/* ------------------------------------------------------------------ */
void tco_bel(const tco_model *m, double *Bel, int i_begin, int i_end) {
  const double B_dx = 1.e-6;
  const long long ne = m->nelems, np = m->np;
#pragma omp parallel
  {
    double *atmp = (double *)malloc(sizeof(double) * 9 * np); /* atmp(3,3,np): [j][ik][c] */
#pragma omp for schedule(dynamic, 100)
    for (int i = i_begin; i < i_end; i++) {
      double pts_i[3][3], evec_i[3][3];
      double area_i = m->ca[i];
      load_cell(m, i, pts_i, evec_i);
      const double *norm_j = m->norm + 3 * i;
      for (long long j = 0; j < np; j++) {
        double pt_j[3] = {m->r[3 * j], m->r[3 * j + 1], m->r[3 * j + 2]};
        double dl_min = 1.e99;
        double dl_max = sqrt(fmax(area_i, m->va[j] / (PI * PI)));
        for (int ii = 0; ii < 3; ii++) {
          double dx = pts_i[ii][0] - pt_j[0], dy = pts_i[ii][1] - pt_j[1], dz = pts_i[ii][2] - pt_j[2];
          double d = sqrt(dx * dx + dy * dy + dz * dz);
          dl_min = fmin(dl_min, d);
          dl_max = fmax(dl_max, d);
        }
        int is_neighbor = (dl_min < 1.e-8);
        int iquad = tco_iquad(dl_min, dl_max);
        double *a = atmp + 9 * j;
        if (iquad > 10) {
          double diffvec[3] = {0.0, 0.0, 0.0};
          if (is_neighbor)
            for (int d = 0; d < 3; d++) pt_j[d] = pt_j[d] - norm_j[d] * 10.0 * B_dx;
          for (int ik = 1; ik <= 2; ik++) {
            if (ik == 2)
              for (int d = 0; d < 3; d++) pt_j[d] = pt_j[d] + norm_j[d] * 20.0 * B_dx;
            for (int jj = 0; jj < 3; jj++) {
              pt_j[jj] = pt_j[jj] + B_dx;
              double tmp = tco_phipot(pts_i, pt_j);
              diffvec[jj] = diffvec[jj] + tmp / (2.0 * B_dx);
              pt_j[jj] = pt_j[jj] - 2.0 * B_dx;
              tmp = tco_phipot(pts_i, pt_j);
              diffvec[jj] = diffvec[jj] - tmp / (2.0 * B_dx);
              pt_j[jj] = pt_j[jj] + B_dx;
            }
            if (!is_neighbor) break;
          }
          if (is_neighbor)
            for (int d = 0; d < 3; d++) diffvec[d] = diffvec[d] / 2.0;
          for (int ik = 0; ik < 3; ik++) {
            a[3 * ik + 0] = diffvec[1] * evec_i[ik][2] - diffvec[2] * evec_i[ik][1];
            a[3 * ik + 1] = diffvec[2] * evec_i[ik][0] - diffvec[0] * evec_i[ik][2];
            a[3 * ik + 2] = diffvec[0] * evec_i[ik][1] - diffvec[1] * evec_i[ik][0];
          }
        } else {
          int nq = TCQ_NP[iquad];
          const double *w = TCQ_WTS + TCQ_OFF[iquad];
          for (int ik = 0; ik < 3; ik++) {
            double diffvec[3] = {0.0, 0.0, 0.0};
            for (int ii = 0; ii < nq; ii++) {
              double x[3], pt_i[3], cr[3];
              quad_point(iquad, ii, pts_i, x);
              for (int d = 0; d < 3; d++) pt_i[d] = pt_j[d] - x[d];
              cross3(evec_i[ik], pt_i, cr);
              double s2 = pt_i[0] * pt_i[0] + pt_i[1] * pt_i[1] + pt_i[2] * pt_i[2];
              double den = pow(s2, 1.5);
              for (int d = 0; d < 3; d++) diffvec[d] = diffvec[d] + cr[d] * w[ii] / den;
            }
            for (int d = 0; d < 3; d++) a[3 * ik + d] = diffvec[d] * area_i;
          }
        }
      }
      for (int ii = 0; ii < 3; ii++) {
        int ik = m->pmap[m->lc[3 * i + ii]];
        if (ik == 0) continue;
        for (long long j = 0; j < np; j++)
          for (int jj = 0; jj < 3; jj++) {
#pragma omp atomic
            Bel[(jj * np + j) * ne + (ik - 1)] += atmp[9 * j + 3 * ii + jj];
          }
      }
      for (int ii = m->kfh[i]; ii < m->kfh[i + 1]; ii++) {
        int ik = abs(m->lfh[2 * ii]) + m->np_active;
        int lv = m->lfh[2 * ii + 1];
        for (long long j = 0; j < np; j++)
          for (int jj = 0; jj < 3; jj++) {
            double v = isign(m->lfh[2 * ii]) * atmp[9 * j + 3 * lv + jj];
#pragma omp atomic
            Bel[(jj * np + j) * ne + (ik - 1)] += v;
          }
      }
    }
    free(atmp);
  }
}

/* Filament Biot-Savart at mesh vertices, thin_wall.F90:2119-2141 / :2147-2168.
 * out Fortran (ld, nsets, 3)-like with caller-chosen strides:
 *   out[jj*stride_c + j*stride_set + i*stride_pt] += ecc(jj)   (no scaling). */
void tco_filament_bfield(const tco_model *m, const tco_coils *cs, double *out, long long stride_pt,
                         long long stride_set, long long stride_c) {
#pragma omp parallel for
  for (int i = 0; i < m->np; i++) {
    const double *pt_j = m->r + 3 * i;
    for (int j = 0; j < cs->nsets; j++) {
      double ecc[3] = {0.0, 0.0, 0.0};
      if (cs->sens_mask && cs->sens_mask[j]) continue;
      for (int k = cs->set_ptr[j]; k < cs->set_ptr[j + 1]; k++) {
        double diffvec[3] = {0.0, 0.0, 0.0};
        for (int kk = cs->fil_ptr[k] + 1; kk < cs->fil_ptr[k + 1]; kk++) {
          const double *a = cs->pts + 3 * kk, *b = cs->pts + 3 * (kk - 1);
          double cvec[3], cpt[3], dv[3], cr[3];
          for (int d = 0; d < 3; d++) {
            cvec[d] = a[d] - b[d];
            cpt[d] = (a[d] + b[d]) / 2.0;
            dv[d] = pt_j[d] - cpt[d];
          }
          cross3(cvec, dv, cr);
          double den = pow(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2], 1.5);
          for (int d = 0; d < 3; d++) diffvec[d] = diffvec[d] + cr[d] / den;
        }
        for (int d = 0; d < 3; d++) ecc[d] = ecc[d] + cs->scales[k] * diffvec[d];
      }
      for (int jj = 0; jj < 3; jj++) out[jj * stride_c + j * stride_set + i * stride_pt] += ecc[jj];
    }
  }
}

/* oft_simple_hash (Jenkins one-at-a-time), src/base/oft_local_c.c:86-98 */
int32_t tco_simple_hash(const uint8_t *key, long length) {
  uint32_t hash = 0;
  for (long i = 0; i < length; i++) {
    hash += key[i];
    hash += hash << 10;
    hash ^= hash >> 6;
  }
  hash += hash << 3;
  hash ^= hash >> 11;
  hash += hash << 15;
  return (int32_t)hash;
}

int tco_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void tco_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* quadrature table accessors (for the KAT test) */
int tco_quad_np(int iquad) { return TCQ_NP[iquad]; }
void tco_quad_get(int iquad, double *pts, double *wts) {
  for (int q = 0; q < TCQ_NP[iquad]; q++) {
    for (int d = 0; d < 3; d++) pts[3 * q + d] = TCQ_PTS[TCQ_OFF[iquad] + q][d];
    wts[q] = TCQ_WTS[TCQ_OFF[iquad] + q];
  }
}

/* Batched wrappers for the kernel-level parity tests (same functions, many inputs). */
void tco_pair_T_batch(int n, const double *Pi, const double *Ai, const double *Pj, const double *Aj,
                      double *T, int *iquad) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int k = 0; k < n; k++)
    T[k] = tco_pair_T((const double(*)[3])(Pi + 9 * (size_t)k), Ai[k], (const double(*)[3])(Pj + 9 * (size_t)k),
                      Aj[k], iquad + k);
}
void tco_phipot_batch(int n, const double *tri, const double *pt, double *out) {
#pragma omp parallel for
  for (int k = 0; k < n; k++) out[k] = tco_phipot((const double(*)[3])(tri + 9 * (size_t)k), pt + 3 * (size_t)k);
}

/* ------------------------------------------------------------------ */
/* Rows of the self-inductance matrix from the per-entry definition     */
/*  -- an independent restatement of what the loop nest  */
/* of thin_wall.F90:1008-1154 accumulates, used to check rows of large  */
/* matrices without running the full O(nc^2) loop:                      */
/*   L[a][b] = 1/(4 pi) sum_{c1 : E_c1[a]!=0} sum_{c2 : E_c2[b]!=0}     */
/*             (E_c1[a].E_c2[b]) T(cell of the smaller DOF analytic)    */
/* dofs[] are 0-based vertex/hole DOF ids; out[nd][nelems].             */
/* ------------------------------------------------------------------ */
static int cell_dofs(const tco_model *m, int c, int *dof, double (*E)[3]) {
  /* distinct DOFs of a cell with their summed basis vectors */
  int n = 0;
  for (int k = 0; k < 3; k++) {
    int d = m->pmap[m->lc[3 * c + k]] - 1;
    if (d < 0) continue;
    int s = 0;
    while (s < n && dof[s] != d) s++;
    if (s == n) { dof[n] = d; E[n][0] = E[n][1] = E[n][2] = 0.0; n++; }
    for (int x = 0; x < 3; x++) E[s][x] += m->qbasis[9 * c + 3 * k + x];
  }
  for (int ii = m->kfh[c]; ii < m->kfh[c + 1]; ii++) {
    int d = m->np_active + abs(m->lfh[2 * ii]) - 1, k = m->lfh[2 * ii + 1];
    int s = 0;
    while (s < n && dof[s] != d) s++;
    if (s == n) { dof[n] = d; E[n][0] = E[n][1] = E[n][2] = 0.0; n++; }
    for (int x = 0; x < 3; x++) E[s][x] += isign(m->lfh[2 * ii]) * m->qbasis[9 * c + 3 * k + x];
  }
  return n;
}

/* absum (optional): A[r][b] = sum of the magnitudes of the terms entry (a_r, b) is summed from.  |L| << A marks entries
 * dominated by cancellation, where the reference's own result moves by ~eps*A with the (atomic, thread-dependent)
 * summation order (thin_wall.F90:1092-1121). */
void tco_lmat_rows2(const tco_model *m, int nd, const int *dofs, double *out, double *absum) {
  const long long N = m->nelems;
  memset(out, 0, sizeof(double) * (size_t)nd * (size_t)N);
  if (absum) memset(absum, 0, sizeof(double) * (size_t)nd * (size_t)N);
  for (int r = 0; r < nd; r++) {
    const int a = dofs[r];
    double *row = out + (size_t)r * N;
    double *arow = absum ? absum + (size_t)r * N : NULL;
    for (int c1 = 0; c1 < m->nc; c1++) {
      int d1[16];
      double E1[16][3];
      int n1 = cell_dofs(m, c1, d1, E1), s1 = -1;
      for (int s = 0; s < n1; s++)
        if (d1[s] == a) s1 = s;
      if (s1 < 0) continue;
      double P1[3][3], Etmp[3][3];
      load_cell(m, c1, P1, Etmp);
#pragma omp parallel for schedule(dynamic, 256)
      for (int c2 = 0; c2 < m->nc; c2++) {
        int d2[16];
        double E2[16][3], P2[3][3], Et2[3][3];
        int n2 = cell_dofs(m, c2, d2, E2);
        if (n2 == 0) continue;
        load_cell(m, c2, P2, Et2);
        double T12 = 0.0, T21 = 0.0;
        int have12 = 0, have21 = 0;
        for (int s = 0; s < n2; s++) {
          const int b = d2[s];
          double T;
          if (a <= b) {
            if (!have12) { T12 = tco_pair_T(P1, m->ca[c1], P2, m->ca[c2], NULL); have12 = 1; }
            T = T12;
          } else {
            if (!have21) { T21 = tco_pair_T(P2, m->ca[c2], P1, m->ca[c1], NULL); have21 = 1; }
            T = T21;
          }
          const double v = dot3(E1[s1], E2[s]) * T / (4.0 * PI);
#pragma omp atomic
          row[b] += v;
          if (arow) {
#pragma omp atomic
            arow[b] += fabs(v);
          }
        }
      }
    }
  }
}

void tco_lmat_rows(const tco_model *m, int nd, const int *dofs, double *out) { tco_lmat_rows2(m, nd, dofs, out, NULL); }

/* ================================================================== */
/* matrix-free apply, HODLR dense-block builders        */
/* ================================================================== */

/* Pair integral of tw_compute_Lmat_MF, thin_wall.F90:1243-1337: the 3-level heuristic
 * (subtended angle > pi/8 or dl_min/dl_max < 0.95 -> order 10, 25 points; angle > pi/4 or ratio
 * < 0.75 -> analytic potential of cell i at the 25 points of cell j; else order 6, 12 points).
 * cls_out: 0 far, 1 close, 2 very close. */
static const double MF_TOLS[2] = {0.75, 0.95}; /* quad_tols(1:2), thin_wall.F90:156 */
double tco_mf_pair_T(const double pts_i[3][3], double area_i, const double pts_j[3][3], double area_j,
                     int *cls_out) {
  int close_flag = 0, vvclose_flag = 0;
  double dl_max = -1.e99, dl_min;
  for (int ii = 0; ii < 3; ii++) {
    double pt_j[3], pt_i[3], tmp;
    for (int d = 0; d < 3; d++) pt_j[d] = pts_j[0][d] - pts_i[ii][d];
    tmp = sqrt(pt_j[0] * pt_j[0] + pt_j[1] * pt_j[1] + pt_j[2] * pt_j[2]);
    if (tmp < 1.e-10) { close_flag = 1; break; }
    for (int d = 0; d < 3; d++) pt_j[d] = pt_j[d] / tmp;
    for (int d = 0; d < 3; d++) pt_i[d] = pts_j[1][d] - pts_i[ii][d];
    tmp = sqrt(pt_i[0] * pt_i[0] + pt_i[1] * pt_i[1] + pt_i[2] * pt_i[2]);
    if (tmp < 1.e-10) { close_flag = 1; break; }
    for (int d = 0; d < 3; d++) pt_i[d] = pt_i[d] / tmp;
    dl_max = fmax(dl_max, fabs(acos(dot3(pt_j, pt_i))));
    for (int d = 0; d < 3; d++) pt_i[d] = pts_j[2][d] - pts_i[ii][d];
    tmp = sqrt(pt_i[0] * pt_i[0] + pt_i[1] * pt_i[1] + pt_i[2] * pt_i[2]);
    if (tmp < 1.e-10) { close_flag = 1; break; }
    for (int d = 0; d < 3; d++) pt_i[d] = pt_i[d] / tmp;
    dl_max = fmax(dl_max, fabs(acos(dot3(pt_j, pt_i))));
  }
  if (dl_max > PI / 8.0) {
    close_flag = 1;
    if (dl_max > PI / 4.0) vvclose_flag = 1;
  } else {
    dl_min = 1.e99;
    dl_max = -1.e99;
    for (int ii = 0; ii < 3; ii++)
      for (int jj = 0; jj < 3; jj++) {
        double dx = pts_i[ii][0] - pts_j[jj][0], dy = pts_i[ii][1] - pts_j[jj][1], dz = pts_i[ii][2] - pts_j[jj][2];
        double d = sqrt(dx * dx + dy * dy + dz * dz);
        dl_min = fmin(dl_min, d);
        dl_max = fmax(dl_max, d);
      }
    if (dl_min / dl_max < MF_TOLS[1]) close_flag = 1;
    if (dl_min / dl_max < MF_TOLS[0]) vvclose_flag = 1;
  }
  double tmp = 0.0;
  if (cls_out) *cls_out = close_flag ? (vvclose_flag ? 2 : 1) : 0;
  if (close_flag && vvclose_flag) {
    const int iq = 10, nq = TCQ_NP[iq];
    const double *w = TCQ_WTS + TCQ_OFF[iq];
    for (int jj = 0; jj < nq; jj++) {
      double pt_j[3];
      quad_point(iq, jj, pts_j, pt_j);
      tmp = tmp + w[jj] * tco_phipot(pts_i, pt_j);
    }
    return tmp * area_j;
  }
  const int iq = close_flag ? 10 : 6, nq = TCQ_NP[iq];
  const double *w = TCQ_WTS + TCQ_OFF[iq];
  for (int ii = 0; ii < nq; ii++) {
    double pt_i[3];
    quad_point(iq, ii, pts_i, pt_i);
    for (int jj = 0; jj < nq; jj++) {
      double pt_j[3];
      quad_point(iq, jj, pts_j, pt_j);
      double dx = pt_i[0] - pt_j[0], dy = pt_i[1] - pt_j[1], dz = pt_i[2] - pt_j[2];
      tmp = tmp + w[jj] * w[ii] / sqrt(dx * dx + dy * dy + dz * dz);
    }
  }
  return tmp * area_i * area_j;
}

/* tw_compute_Lmat_MF, thin_wall.F90:1190-1414: b(:,irhs) = M a(:,irhs), a Fortran (row%nelems,nrhs),
 * b Fortran (col%nelems,nrhs); V-coil parts are not computed by the reference either (:1385).
 * counts[3] (optional): pairs per class. */
void tco_lmat_mf(const tco_model *row, const tco_model *col, int nrhs, const double *a, double *b,
                 long long *counts) {
  const long long nr = row->nelems, ncl = col->nelems;
  memset(b, 0, sizeof(double) * (size_t)ncl * nrhs);
  long long c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel for schedule(dynamic, 100) reduction(+ : c0, c1, c2)
  for (int i = 0; i < row->nc; i++) {
    double pts_i[3][3], evec_i[3][3], pts_j[3][3], evec_j[3][3];
    load_cell(row, i, pts_i, evec_i);
    for (int j = 0; j < col->nc; j++) {
      load_cell(col, j, pts_j, evec_j);
      int cls;
      const double tmp = tco_mf_pair_T(pts_i, row->ca[i], pts_j, col->ca[j], &cls);
      if (cls == 0) c0++; else if (cls == 1) c1++; else c2++;
      for (int ii = 0; ii < 3; ii++) {
        int ik = row->pmap[row->lc[3 * i + ii]];
        if (ik == 0) continue;
        for (int jj = 0; jj < 3; jj++) {
          int jk = col->pmap[col->lc[3 * j + jj]];
          if (jk == 0) continue;
          double v = dot3(evec_i[ii], evec_j[jj]) * tmp;
          for (int q = 0; q < nrhs; q++) {
#pragma omp atomic
            b[q * ncl + (jk - 1)] += v * a[q * nr + (ik - 1)];
          }
        }
        for (int jj = col->kfh[j]; jj < col->kfh[j + 1]; jj++) {
          int jk = abs(col->lfh[2 * jj]) + col->np_active;
          double v = isign(col->lfh[2 * jj]) * dot3(evec_i[ii], evec_j[col->lfh[2 * jj + 1]]) * tmp;
          for (int q = 0; q < nrhs; q++) {
#pragma omp atomic
            b[q * ncl + (jk - 1)] += v * a[q * nr + (ik - 1)];
          }
        }
      }
      for (int ii = row->kfh[i]; ii < row->kfh[i + 1]; ii++) {
        int ik = abs(row->lfh[2 * ii]) + row->np_active;
        const double *ei = evec_i[row->lfh[2 * ii + 1]];
        int si = isign(row->lfh[2 * ii]);
        for (int jj = 0; jj < 3; jj++) {
          int jk = col->pmap[col->lc[3 * j + jj]];
          if (jk == 0) continue;
          double v = si * dot3(ei, evec_j[jj]) * tmp;
          for (int q = 0; q < nrhs; q++) {
#pragma omp atomic
            b[q * ncl + (jk - 1)] += v * a[q * nr + (ik - 1)];
          }
        }
        for (int jj = col->kfh[j]; jj < col->kfh[j + 1]; jj++) {
          int jk = abs(col->lfh[2 * jj]) + col->np_active;
          double v = si * isign(col->lfh[2 * jj]) * dot3(ei, evec_j[col->lfh[2 * jj + 1]]) * tmp;
          for (int q = 0; q < nrhs; q++) {
#pragma omp atomic
            b[q * ncl + (jk - 1)] += v * a[q * nr + (ik - 1)];
          }
        }
      }
    }
  }
  for (long long k = 0; k < ncl * nrhs; k++) b[k] = b[k] / (4.0 * PI);
  if (counts) { counts[0] = c0; counts[1] = c1; counts[2] = c2; }
}

/* tw_compute_Lmatblock, thin_wall_hodlr.F90:289-404: dense block over cell sub-lists.  The ROW block's cell is always
 * the analytic side of a near pair, every listed pair is visited (no jmax<imin skip), vertex DOFs only.
 * row_cells/col_cells 0-based; row_inv/col_inv [np] = 1-based index in the block or 0 (oft_tw_block%inv_map).
 * Lmat Fortran (ncol_out,nrow_out): Lmat[(ik-1)*ncol_out + (jk-1)]. */
void tco_lmat_block(const tco_model *row, const tco_model *col, int nrc, const int *row_cells, const int *row_inv,
                    int ncc, const int *col_cells, const int *col_inv, int nrow_out, int ncol_out, double *Lmat) {
  memset(Lmat, 0, sizeof(double) * (size_t)nrow_out * ncol_out);
  for (int irow = 0; irow < nrc; irow++) {
    const int i = row_cells[irow];
    double pts_i[3][3], evec_i[3][3];
    load_cell(row, i, pts_i, evec_i);
#pragma omp parallel for schedule(dynamic, 64)
    for (int jcol = 0; jcol < ncc; jcol++) {
      const int j = col_cells[jcol];
      double pts_j[3][3], evec_j[3][3];
      load_cell(col, j, pts_j, evec_j);
      const double tmp = tco_pair_T(pts_i, row->ca[i], pts_j, col->ca[j], NULL);
      for (int ii = 0; ii < 3; ii++) {
        int ik = row_inv[row->lc[3 * i + ii]];
        if (ik == 0) continue;
        for (int jj = 0; jj < 3; jj++) {
          int jk = col_inv[col->lc[3 * j + jj]];
          if (jk == 0) continue;
          double v = dot3(evec_i[ii], evec_j[jj]) * tmp;
#pragma omp atomic
          Lmat[(size_t)(ik - 1) * ncol_out + (jk - 1)] += v;
        }
      }
    }
  }
  for (size_t k = 0; k < (size_t)nrow_out * ncol_out; k++) Lmat[k] = Lmat[k] / (4.0 * PI);
}

/* tw_compute_LmatHole, thin_wall_hodlr.F90:136-285: columns of L for the hole (and V-coil) DOFs of the row model.
 * Lmat Fortran (col%nelems, row%nholes + row%n_vcoils): Lmat[ik*col%nelems + jk]; Ael2coil Fortran (nelems,n_vcoils),
 * Acoil2coil (n_vcoils,n_vcoils), NULL without V-coils. */
void tco_lmat_hole(const tco_model *row, const tco_model *col, const double *Ael2coil, const double *Acoil2coil,
                   double *Lmat) {
  const long long ncl = col->nelems;
  const int nk = row->nholes + row->n_vcoils;
  memset(Lmat, 0, sizeof(double) * (size_t)ncl * nk);
#pragma omp parallel for schedule(dynamic, 100)
  for (int i = 0; i < row->nc; i++) {
    if (row->kfh[i + 1] - row->kfh[i] == 0) continue;
    double pts_i[3][3], evec_i[3][3], pts_j[3][3], evec_j[3][3];
    load_cell(row, i, pts_i, evec_i);
    for (int j = 0; j < col->nc; j++) {
      load_cell(col, j, pts_j, evec_j);
      const double tmp = tco_pair_T(pts_i, row->ca[i], pts_j, col->ca[j], NULL);
      for (int ii = row->kfh[i]; ii < row->kfh[i + 1]; ii++) {
        const int ik = abs(row->lfh[2 * ii]);
        const double *ei = evec_i[row->lfh[2 * ii + 1]];
        const int si = isign(row->lfh[2 * ii]);
        for (int jj = 0; jj < 3; jj++) {
          int jk = col->pmap[col->lc[3 * j + jj]];
          if (jk <= 0) continue;
          double v = si * dot3(ei, evec_j[jj]) * tmp;
#pragma omp atomic
          Lmat[(size_t)(ik - 1) * ncl + (jk - 1)] += v;
        }
        for (int jj = col->kfh[j]; jj < col->kfh[j + 1]; jj++) {
          int jk = abs(col->lfh[2 * jj]) + col->np_active;
          double v = si * isign(col->lfh[2 * jj]) * dot3(ei, evec_j[col->lfh[2 * jj + 1]]) * tmp;
#pragma omp atomic
          Lmat[(size_t)(ik - 1) * ncl + (jk - 1)] += v;
        }
      }
    }
  }
  if (Ael2coil && Acoil2coil) { /* :253-276 (row == col in every call site) */
    const long long ne = row->nelems;
    for (int i = 0; i < row->np_active + row->nholes; i++)
      for (int j = 0; j < col->n_vcoils; j++) Lmat[(size_t)(col->nholes + j) * ncl + i] = Ael2coil[(size_t)j * ne + i];
    for (int i = 0; i < row->nholes; i++)
      for (int j = 0; j < col->n_vcoils; j++)
        Lmat[(size_t)i * ncl + (row->np_active + col->nholes + j)] = Ael2coil[(size_t)j * ne + row->np_active + i];
    for (int i = 0; i < row->n_vcoils; i++)
      for (int j = 0; j < col->n_vcoils; j++)
        Lmat[(size_t)(row->nholes + i) * ncl + (col->np_active + col->nholes + j)] = Acoil2coil[(size_t)j * row->n_vcoils + i];
  }
  for (size_t k = 0; k < (size_t)ncl * nk; k++) Lmat[k] = Lmat[k] / (4.0 * PI);
}

/* tw_compute_Bops_block, thin_wall_hodlr.F90:580-691: one Cartesian component (dir = 0,1,2) of the B operator for the
 * vertex DOFs of a row block (cells row_cells, inv_map row_inv) at the mesh vertices col_pts (0-based).
 * Bop Fortran (ncp, nrow_out): Bop[(ik-1)*ncp + jcol]. */
void tco_bops_block(const tco_model *m, int nrc, const int *row_cells, const int *row_inv, int nrow_out, int ncp,
                    const int *col_pts, int dir, double *Bop) {
  const double B_dx = 1.e-6;
  memset(Bop, 0, sizeof(double) * (size_t)ncp * nrow_out);
  for (int irow = 0; irow < nrc; irow++) {
    const int i = row_cells[irow];
    double pts_i[3][3], evec_i[3][3];
    const double area_i = m->ca[i];
    load_cell(m, i, pts_i, evec_i);
    const double *norm_j = m->norm + 3 * i;
#pragma omp parallel for schedule(dynamic, 64)
    for (int jcol = 0; jcol < ncp; jcol++) {
      const int j = col_pts[jcol];
      double pt_j[3] = {m->r[3 * j], m->r[3 * j + 1], m->r[3 * j + 2]};
      double dl_min = 1.e99;
      double dl_max = sqrt(fmax(area_i, m->va[j] / (PI * PI)));
      for (int ii = 0; ii < 3; ii++) {
        double dx = pts_i[ii][0] - pt_j[0], dy = pts_i[ii][1] - pt_j[1], dz = pts_i[ii][2] - pt_j[2];
        double d = sqrt(dx * dx + dy * dy + dz * dz);
        dl_min = fmin(dl_min, d);
        dl_max = fmax(dl_max, d);
      }
      const int is_neighbor = (dl_min < 1.e-8);
      const int iquad = tco_iquad(dl_min, dl_max);
      double a[9];
      if (iquad > 10) {
        double diffvec[3] = {0.0, 0.0, 0.0};
        if (is_neighbor)
          for (int d = 0; d < 3; d++) pt_j[d] = pt_j[d] - norm_j[d] * 10.0 * B_dx;
        for (int ik = 1; ik <= 2; ik++) {
          if (ik == 2)
            for (int d = 0; d < 3; d++) pt_j[d] = pt_j[d] + norm_j[d] * 20.0 * B_dx;
          for (int jj = 0; jj < 3; jj++) {
            pt_j[jj] = pt_j[jj] + B_dx;
            double tmp = tco_phipot(pts_i, pt_j);
            diffvec[jj] = diffvec[jj] + tmp / (2.0 * B_dx);
            pt_j[jj] = pt_j[jj] - 2.0 * B_dx;
            tmp = tco_phipot(pts_i, pt_j);
            diffvec[jj] = diffvec[jj] - tmp / (2.0 * B_dx);
            pt_j[jj] = pt_j[jj] + B_dx;
          }
          if (!is_neighbor) break;
        }
        if (is_neighbor)
          for (int d = 0; d < 3; d++) diffvec[d] = diffvec[d] / 2.0;
        for (int ik = 0; ik < 3; ik++) {
          a[3 * ik + 0] = diffvec[1] * evec_i[ik][2] - diffvec[2] * evec_i[ik][1];
          a[3 * ik + 1] = diffvec[2] * evec_i[ik][0] - diffvec[0] * evec_i[ik][2];
          a[3 * ik + 2] = diffvec[0] * evec_i[ik][1] - diffvec[1] * evec_i[ik][0];
        }
      } else {
        const int nq = TCQ_NP[iquad];
        const double *w = TCQ_WTS + TCQ_OFF[iquad];
        for (int ik = 0; ik < 3; ik++) {
          double diffvec[3] = {0.0, 0.0, 0.0};
          for (int ii = 0; ii < nq; ii++) {
            double x[3], pt_i[3], cr[3];
            quad_point(iquad, ii, pts_i, x);
            for (int d = 0; d < 3; d++) pt_i[d] = pt_j[d] - x[d];
            cross3(evec_i[ik], pt_i, cr);
            double s2 = pt_i[0] * pt_i[0] + pt_i[1] * pt_i[1] + pt_i[2] * pt_i[2];
            double den = pow(s2, 1.5);
            for (int d = 0; d < 3; d++) diffvec[d] = diffvec[d] + cr[d] * w[ii] / den;
          }
          for (int d = 0; d < 3; d++) a[3 * ik + d] = diffvec[d] * area_i;
        }
      }
      for (int ii = 0; ii < 3; ii++) {
        int ik = row_inv[m->lc[3 * i + ii]];
        if (ik == 0) continue;
        Bop[(size_t)(ik - 1) * ncp + jcol] += a[3 * ii + dir]; /* distinct jcol per thread, serial over irow */
      }
    }
  }
  for (size_t k = 0; k < (size_t)ncp * nrow_out; k++) Bop[k] = Bop[k] / (4.0 * PI);
}
