/*
 * verify.c -- --verify 1: run the selected GPU time stepper once, gather the interior on rank 0 and
 * compare it with serial reference kernels on the undecomposed global domain.
 * Follows src/verification.c:26-51 (verify), :52-312 (verify_serial_generic), :315-479 and :786-819
 * (std_kernel_*), :823-860 (compare_results_std), :914-952 (verification_printing).
 *
 * The reference kernels here are the product's own verifier (the reference ships one too); they are
 * not the test oracle under oracle/, which is never linked into this binary.  The loops may run
 * OpenMP-parallel over z: each point is a single expression of the previous level, so the bits do
 * not depend on the schedule (the file is built with -ffp-contract=off, the counterpart of the
 * reference building verification.c at -O0, Makefile:4,51-52).
 * The pass criterion is the reference's: the L1 norm of the difference must be exactly zero
 * (with --gpu-contract 1: relative L-infinity <= 1e-12 fp64 / 1e-5 fp32, see verify_compute).
 * The relative L-infinity error (the north star's tolerance: 1e-12 fp64 / 1e-5 fp32) is printed too.
 */
#define _POSIX_C_SOURCE 200112L
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "girih_host.h"

#define IDX(i, j, k) (((uint64_t)(k) * nny + (j)) * nnx + (i))

/* one serial step u <- f(v) over the interior of the global domain */
static void std_kernel(const Parameters *p, const int shape[3], const real_t *coef, real_t *u, const real_t *v,
                       const real_t *roc2) {
  const int nnx = shape[0], nny = shape[1], nnz = shape[2], r = p->stencil.r;
  const uint64_t n = (uint64_t)nnx * nny * nnz;
  const int64_t sx = 1, sy = nnx, sz = (int64_t)nnx * nny;
  const int kind = p->stencil.coeff, box = (p->stencil.shape == GIRIH_BOX), to2 = (p->stencil.time_order == 2);
  int k;
#pragma omp parallel for schedule(static)
  for (k = r; k < nnz - r; k++) {
    int i, j, m;
    for (j = r; j < nny - r; j++)
      for (i = r; i < nnx - r; i++) {
        const uint64_t c = IDX(i, j, k);
        const real_t *vc = v + c;
        real_t acc;
        if (box) { /* src/verification.c:798-814 */
          acc = coef[0] * vc[0] + coef[1] * (vc[sx] + vc[-sx]) + coef[1] * (vc[sy] + vc[-sy]) + coef[1] * (vc[sz] + vc[-sz])
              + coef[2] * (vc[sx - sz] + vc[-sx - sz]) + coef[2] * (vc[sy - sz] + vc[-sy - sz])
              + coef[2] * (vc[sx + sy] + vc[-sx - sy]) + coef[2] * (vc[sx - sy] + vc[-sx + sy])
              + coef[2] * (vc[sx + sz] + vc[-sx + sz]) + coef[2] * (vc[sy + sz] + vc[-sy + sz])
              + coef[3] * (vc[sx + sy + sz] + vc[-sx - sy - sz]) + coef[3] * (vc[sx - sy + sz] + vc[-sx + sy - sz])
              + coef[3] * (vc[-sx - sy + sz] + vc[sx + sy - sz]) + coef[3] * (vc[-sx + sy + sz] + vc[sx - sy - sz]);
        } else if (kind == GIRIH_COEF_VARIABLE_NOSYM) { /* :468-474 */
          acc = coef[c] * vc[0] + coef[c + n] * vc[-sx] + coef[c + 2 * n] * vc[sx] + coef[c + 3 * n] * vc[-sy]
              + coef[c + 4 * n] * vc[sy] + coef[c + 5 * n] * vc[-sz] + coef[c + 6 * n] * vc[sz];
        } else { /* symmetric star: distance m = 1..r, axes x, y, z inside each m (:330-342, :367-370, :437-449) */
          acc = (kind == GIRIH_COEF_CONSTANT ? coef[0] : coef[c]) * vc[0];
          for (m = 1; m <= r; m++) {
            real_t cx, cy, cz;
            if (kind == GIRIH_COEF_CONSTANT) cx = cy = cz = coef[m];
            else if (kind == GIRIH_COEF_VARIABLE) cx = cy = cz = coef[c + (uint64_t)m * n];
            else { cx = coef[c + (uint64_t)(1 + 3 * (m - 1)) * n]; cy = coef[c + (uint64_t)(2 + 3 * (m - 1)) * n]; cz = coef[c + (uint64_t)(3 + 3 * (m - 1)) * n]; }
            acc = acc + cx * (vc[m * sx] + vc[-m * sx]);
            acc = acc + cy * (vc[m * sy] + vc[-m * sy]);
            acc = acc + cz * (vc[m * sz] + vc[-m * sz]);
          }
        }
        if (to2) u[c] = ((real_t)2.0) * vc[0] - u[c] + roc2[c] * acc; /* :344-348 */
        else u[c] = acc;
      }
  }
}

static void *xalloc(size_t bytes) {
  void *ptr = NULL;
  if (posix_memalign(&ptr, 64, bytes ? bytes : 1) != 0) girih_fatal(NULL, "no sufficient memory");
  return ptr;
}

/* reference solution of the global problem after p->nt steps of the reference loop; returns u */
static real_t *serial_reference(const Parameters *p, int shape[3]) {
  const int r = p->stencil.r;
  const int nnx = p->stencil_shape[0] + 2 * r, nny = p->stencil_shape[1] + 2 * r, nnz = p->stencil_shape[2] + 2 * r;
  const uint64_t n = (uint64_t)nnx * nny * nnz;
  real_t *u = (real_t *)xalloc(sizeof(real_t) * n), *v = (real_t *)xalloc(sizeof(real_t) * n), *roc2 = NULL, *coef;
  uint64_t i, m, ax, csize;
  int x, y, z, it;
  shape[0] = nnx; shape[1] = nny; shape[2] = nnz;
  if (p->stencil.time_order == 2) roc2 = (real_t *)xalloc(sizeof(real_t) * n);
  switch (p->stencil.coeff) { /* src/verification.c:92-197 */
    case GIRIH_COEF_CONSTANT:
      coef = (real_t *)xalloc(sizeof(real_t) * 11);
      memset(coef, 0, sizeof(real_t) * 11);   /* box: coef[2..3] stay zero, see arrays_allocate() */
      for (i = 0; i < (uint64_t)r + 1; i++) coef[i] = p->g_coef[i];
      break;
    case GIRIH_COEF_VARIABLE:
      csize = n * (uint64_t)(1 + r);
      coef = (real_t *)xalloc(sizeof(real_t) * csize);
      for (m = 0; m <= (uint64_t)r; m++)
        for (i = 0; i < n; i++) coef[i + m * n] = p->g_coef[m];
      break;
    case GIRIH_COEF_VARIABLE_AXSYM:
      csize = n * (uint64_t)(1 + 3 * r);
      coef = (real_t *)xalloc(sizeof(real_t) * csize);
      for (i = 0; i < n; i++) coef[i] = p->g_coef[0];
      for (m = 0; m < (uint64_t)r; m++)
        for (ax = 0; ax < 3; ax++)
          for (i = 0; i < n; i++) coef[i + n + 3 * m * n + ax * n] = p->g_coef[m + 1];
      break;
    default:
      csize = n * (uint64_t)(1 + 6 * r);
      coef = (real_t *)xalloc(sizeof(real_t) * csize);
      for (i = 0; i < n; i++) coef[i] = p->g_coef[0];
      for (m = 0; m < (uint64_t)r; m++)
        for (ax = 0; ax < 3; ax++)
          for (i = 0; i < n; i++) {
            coef[i + n + 6 * m * n + 2 * ax * n] = p->g_coef[m + 1];
            coef[i + n + 6 * m * n + (2 * ax + 1) * n] = p->g_coef[m + 1];
          }
  }
  for (i = 0; i < n; i++) { /* :223-228 */
    u[i] = 0.0;
    v[i] = 0.0;
    if (roc2) roc2[i] = 0.0;
  }
  for (z = r; z < p->stencil_shape[2] + r; z++) /* :230-241 */
    for (y = r; y < p->stencil_shape[1] + r; y++)
      for (x = r; x < p->stencil_shape[0] + r; x++) {
        const real_t w = 1.0 / 3 * (1.0 * x / p->stencil_shape[0] + 1.0 * y / p->stencil_shape[1] + 1.0 * z / p->stencil_shape[2]);
        u[IDX(x, y, z)] = w * 1.845703;
        v[IDX(x, y, z)] = w * 1.845703;
        if (roc2) roc2[IDX(x, y, z)] = w * 1.845703;
      }
  for (z = 0; z < nnz; z++) /* :243-250 */
    for (y = 0; y < nny; y++) {
      u[IDX(0, y, z)] += BOUNDARY_SRC_VAL;
      v[IDX(0, y, z)] += BOUNDARY_SRC_VAL;
      u[IDX(nnx - 1, y, z)] += BOUNDARY_SRC_VAL;
      v[IDX(nnx - 1, y, z)] += BOUNDARY_SRC_VAL;
    }
  for (it = 0; it < p->nt; it += 2) { /* :281-284 */
    std_kernel(p, shape, coef, u, v, roc2);
    std_kernel(p, shape, coef, v, u, roc2);
  }
  free(v);
  free(coef);
  if (roc2) free(roc2);
  return u;
}

static void verification_printing(const Parameters *p) { /* src/verification.c:914-952 */
  const char *coeff_type;
  if (p->mpi_rank != 0) return;
  switch (p->stencil.coeff) {
    case GIRIH_COEF_CONSTANT: coeff_type = "const    "; break;
    case GIRIH_COEF_VARIABLE: coeff_type = "var      "; break;
    case GIRIH_COEF_VARIABLE_AXSYM: coeff_type = "var_axsym"; break;
    case GIRIH_COEF_VARIABLE_NOSYM: coeff_type = "var_nosym"; break;
    default: coeff_type = "Solar kernel";
  }
  printf("#ts:%s stencil:%s|R:%d|T:%d|%s nt:%03d thrd:%d prec.:%s %s dom:(%d,%03d,%03d) top:(%d,%d,%d) ",
         TSList[p->target_ts].name, p->stencil.name, p->stencil.r, p->stencil.time_order, coeff_type, p->nt,
         p->num_threads, (sizeof(real_t) == 4) ? "SP" : "DP", (p->halo_concat == 0) ? "no-concat" : "   concat",
         p->lstencil_shape[0], p->lstencil_shape[1], p->lstencil_shape[2], p->t.shape[0], p->t.shape[1], p->t.shape[2]);
  if (p->target_ts == 2) {
    printf("TB:%d wf:%d ", p->t_dim, p->wavefront);
    printf("|thrd_group|:%d ", p->thread_group_size);
    printf("num-wf:%d ", p->num_wf);
  }
  if (p->verbose == 1) print_param(p);
}

/* runs the stepper and the comparison; returns 0 on PASS (collective over the ranks) */
int verify_compute(Parameters *p, double *max_err, double *l1_err, double *max_ref) {
  const int r = p->stencil.r;
  const int nx = p->stencil_shape[0], ny = p->stencil_shape[1], nz = p->stencil_shape[2];
  real_t *aggr;
  int i, j, k, rc, broken = 0;

  if (is_solar(p)) return solar_verify_compute(p, max_err, l1_err, max_ref);
  arrays_allocate(p);
  init_coeff(p);
  domain_data_fill(p);
  gpu_attach(p);
  TSList[p->target_ts].func(p);
  rc = girih_gpu_download(p->gpu, p->U1, NULL);
  if (rc != GIRIH_OK) girih_fatal(NULL, "girih_gpu_download: %s", girih_gpu_strerror(rc));

  /* aggregate the sub-domain interiors (src/verification.c:955-1040) */
  aggr = (real_t *)team_shared_alloc(sizeof(real_t) * p->n_stencils, p->mpi_rank);
  for (k = 0; k < p->lstencil_shape[2]; k++)
    for (j = 0; j < p->lstencil_shape[1]; j++)
      for (i = 0; i < p->lstencil_shape[0]; i++)
        aggr[((uint64_t)(k + p->gb[2]) * ny + (j + p->gb[1])) * nx + (i + p->gb[0])] =
            p->U1[((uint64_t)(k + r) * p->ldomain_shape[1] + (j + r)) * p->ldomain_shape[0] + (i + r)];
  team_barrier();

  if (p->mpi_rank == 0) {
    int shape[3];
    real_t *u = serial_reference(p, shape);
    const int nnx = shape[0], nny = shape[1];
    real_t diff_l1 = 0.0, maxe = 0.0; /* accumulated in real_t like src/verification.c:829 */
    double mref = 0.0;
    for (k = 0; k < nz; k++)
      for (j = 0; j < ny; j++)
        for (i = 0; i < nx; i++) {
          const real_t a = u[IDX(i + r, j + r, k + r)];
          const real_t d = fabs(a - aggr[((uint64_t)k * ny + j) * nx + i]);
          if (d > maxe) maxe = d;
          diff_l1 += d;
          if (fabs((double)a) > mref) mref = fabs((double)a);
        }
    broken = (diff_l1 > 0.0) || (diff_l1 * 0 != 0) || (diff_l1 != diff_l1);
    if (p->gpu_contract) {
      /* FMA-contracted arithmetic rounds differently from this serial verifier (as the reference's own
       * -xHost/-mfma kernels do from its -O0 verifier): the criterion is the north star's tolerance */
      const double rel = mref > 0 ? (double)maxe / mref : (double)maxe;
      broken = !(rel <= (sizeof(real_t) == 8 ? 1e-12 : 1e-5));
    }
    *max_err = maxe; *l1_err = diff_l1; *max_ref = mref;
    free(u);
  }
  team_shared_free(aggr, p->mpi_rank);
  gpu_detach(p);
  arrays_free(p);
  return broken;
}

void verify(Parameters *p) {
  double maxe = 0, l1 = 0, mref = 0;
  int broken;
  verification_printing(p);
  broken = verify_compute(p, &maxe, &l1, &mref);
  if (p->mpi_rank != 0) return;
  if (broken) { /* src/verification.c:842-849 */
    printf("Max snapshot abs. err.:%e  L1 norm:%e\n", maxe, l1);
    printf("relative Linf error: %e (max |ref| %e)\n", mref > 0 ? maxe / mref : maxe, mref);
    fprintf(stderr, "BROKEN KERNEL\n");
    exit(1);
  }
  printf("eMax:%.3e|eL1:%.3e", maxe, l1);
  printf("-PASSED\n");
  if (p->verbose == 1) printf("relative Linf error: %e (max |ref| %e)\n", mref > 0 ? maxe / mref : maxe, mref);
}
