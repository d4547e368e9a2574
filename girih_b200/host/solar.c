/*
 * solar.c -- host side of table slot 6, "solar" (12 complex field components, 28 complex coefficient arrays):
 * allocation sizes, coefficient and field initialisation (src/utils.c:168-172, 199-201, 483-489, 698-727), and the
 * serial verifier of --verify 1 (src/verification.c:253-268 fill, :481-784 reference kernels, :861-906 comparison).
 *
 * The verifier kernels below are the product's own CPU reference, as the reference ships one; they are not the test
 * oracle under oracle/, which is never linked into this binary.  Built with -ffp-contract=off (the counterpart of
 * the reference building verification.c at -O0).
 */
#define _POSIX_C_SOURCE 200112L
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "girih_host.h"

int is_solar(const Parameters *p) { return p->stencil.coeff == GIRIH_COEF_SOLAR; }

/* src/utils.c:483-489: the value is picked by the flat index modulo 10 */
void solar_init_coeff(Parameters *p) {
  const uint64_t n = p->ln_domain * 28u * 2u;
  uint64_t i;
  for (i = 0; i < n; i++) p->coef[i] = p->g_coef[i % 10];
}

/* fills u[12][nnz][nny][nnx][2] for a (sub)domain whose first cell is global cell gb; gs = global interior shape.
 * src/utils.c:698-727 (and src/verification.c:253-268 with gb = 0): every cell, frame included */
static void solar_fill(real_t *u, const int shape[3], const int gs[3], const int gb[3]) {
  const uint64_t n = (uint64_t)shape[0] * shape[1] * shape[2];
  int f, i, j, k;
  for (f = 0; f < 12; f++)
    for (k = 0; k < shape[2]; k++)
      for (j = 0; j < shape[1]; j++)
        for (i = 0; i < shape[0]; i++) {
          const uint64_t gi = (uint64_t)i + gb[0], gj = (uint64_t)j + gb[1], gk = (uint64_t)k + gb[2];
          const real_t w = 1.0 / (3.0) * (1.0 * gi / gs[0] + 1.0 * gj / gs[1] + 1.0 * gk / gs[2]);
          const uint64_t idx = 2 * ((((uint64_t)k * shape[1] + j) * shape[0] + i) + n * f);
          u[idx] = w * 1.845703;
          u[idx + 1] = w * 1.845703 / 3.0;
        }
}

void solar_domain_fill(Parameters *p) { solar_fill(p->U1, p->ldomain_shape, p->stencil_shape, p->gb); }

/* ---- serial reference: one time step = H update then E update of every interior cell ---------------------- */
typedef struct { int own, c, t, bnd, p, q, axis, form; } solar_comp;
/* own field, coefficient arrays (c, t, boundary source or -1), the two source fields, the axis of the staggered
 * difference (0 x, 1 y, 2 z) and the order of its four terms:
 *   0: P[i]-P[s]+Q[i]-Q[s]   1: P[s]-P[i]+Q[s]-Q[i]   2: P[s]+Q[s]-P[i]-Q[i]   3: P[i]+Q[i]-P[s]-Q[s]
 * H components look at the neighbour below (s = i - stride), E components at the one above. */
static const solar_comp SOLAR_H[6] = {
    {0, 0, 6, 13, 10, 6, 2, 0},  /* Hy_x  src/verification.c:531-537 */
    {1, 1, 7, -1, 10, 6, 1, 1},  /* Hz_x  :548-554 */
    {2, 2, 8, 12, 8, 7, 2, 1},   /* Hx_y  :565-571 */
    {3, 3, 9, -1, 8, 7, 0, 0},   /* Hz_y  :582-588 */
    {4, 4, 10, -1, 9, 11, 1, 0}, /* Hx_z  :599-605 */
    {5, 5, 11, -1, 9, 11, 0, 2}, /* Hy_z  :616-622 */
};
static const solar_comp SOLAR_E[6] = {
    {6, 14, 20, -1, 1, 3, 1, 1},  /* Ex_z  :680-686 */
    {7, 15, 21, -1, 1, 3, 0, 3},  /* Ey_z  :697-703 */
    {8, 16, 22, 27, 2, 4, 2, 1},  /* Ey_x  :714-720 */
    {9, 17, 23, -1, 2, 4, 1, 3},  /* Ez_x  :731-737 */
    {10, 18, 24, 26, 0, 5, 2, 0}, /* Ex_y  :748-754 */
    {11, 19, 25, -1, 0, 5, 0, 1}, /* Ez_y  :765-771 */
};

static real_t stag(int form, real_t pi, real_t ps, real_t qi, real_t qs) {
  switch (form) {
    case 0: return pi - ps + qi - qs;
    case 1: return ps - pi + qs - qi;
    case 2: return ps + qs - pi - qi;
    default: return pi + qi - ps - qs;
  }
}

static void solar_phase(const int shape[3], const real_t *coef, real_t *u, const solar_comp *comps, int is_h) {
  const int nnx = shape[0], nny = shape[1], nnz = shape[2];
  const uint64_t ln2 = 2 * (uint64_t)nnx * nny * nnz;
  const int64_t stride[3] = {2, 2 * (int64_t)nnx, 2 * (int64_t)nnx * nny};
  int k;
#pragma omp parallel for schedule(static)
  for (k = 1; k < nnz - 1; k++) {
    int j, x, m;
    for (j = 1; j < nny - 1; j++)
      for (x = 1; x < nnx - 1; x++) {
        const uint64_t i = 2 * (((uint64_t)k * nny + j) * nnx + x);
        for (m = 0; m < 6; m++) {
          const solar_comp *sc = &comps[m];
          real_t *f = u + ln2 * sc->own;
          const real_t *P = u + ln2 * sc->p, *Q = u + ln2 * sc->q, *c = coef + ln2 * sc->c, *t = coef + ln2 * sc->t;
          const uint64_t s = is_h ? i - stride[sc->axis] : i + stride[sc->axis];
          const real_t dR = stag(sc->form, P[i], P[s], Q[i], Q[s]);
          const real_t dI = stag(sc->form, P[i + 1], P[s + 1], Q[i + 1], Q[s + 1]);
          real_t re = f[i] * t[i] - f[i + 1] * t[i + 1], im = f[i] * t[i + 1] + f[i + 1] * t[i];
          if (sc->bnd >= 0) {
            re = re + coef[ln2 * sc->bnd + i];
            im = im + coef[ln2 * sc->bnd + i + 1];
          }
          if (is_h) {
            re = re - c[i] * dR + c[i + 1] * dI;
            im = im - c[i] * dI - c[i + 1] * dR;
          } else {
            re = re + c[i] * dR - c[i + 1] * dI;
            im = im + c[i] * dI + c[i + 1] * dR;
          }
          f[i] = re;
          f[i + 1] = im;
        }
      }
  }
}

static void *xalloc(size_t bytes) {
  void *ptr = NULL;
  if (posix_memalign(&ptr, 64, bytes ? bytes : 1) != 0) girih_fatal(NULL, "no sufficient memory");
  return ptr;
}

/* --verify 1 for the solar slot: stepper on the GPU, serial reference on the host, every real of the 12 fields compared
 * (the reference looks at the real parts, src/verification.c:873-880; the imaginary parts are compared here as well) */
int solar_verify_compute(Parameters *p, double *max_err, double *l1_err, double *max_ref) {
  const int shape[3] = {p->stencil_shape[0] + 2, p->stencil_shape[1] + 2, p->stencil_shape[2] + 2};
  const uint64_t n24 = 24 * (uint64_t)shape[0] * shape[1] * shape[2];
  const int zero[3] = {0, 0, 0};
  real_t *u, *coef, diff_l1 = 0.0, maxe = 0.0;
  double mref = 0.0;
  uint64_t i;
  int it, rc;

  arrays_allocate(p);
  init_coeff(p);
  domain_data_fill(p);
  gpu_attach(p);
  TSList[p->target_ts].func(p);
  rc = girih_gpu_download(p->gpu, p->U1, NULL);
  if (rc != GIRIH_OK) girih_fatal(NULL, "girih_gpu_download: %s", girih_gpu_strerror(rc));

  u = (real_t *)xalloc(sizeof(real_t) * n24);
  coef = (real_t *)xalloc(sizeof(real_t) * (n24 / 24) * 56);
  for (i = 0; i < (n24 / 24) * 56; i++) coef[i] = p->g_coef[i % 10]; /* src/verification.c:200-208 */
  solar_fill(u, shape, p->stencil_shape, zero);
  for (it = 0; it < p->nt; it += 2) { /* :281-284: two time steps per iteration */
    solar_phase(shape, coef, u, SOLAR_H, 1);
    solar_phase(shape, coef, u, SOLAR_E, 0);
    solar_phase(shape, coef, u, SOLAR_H, 1);
    solar_phase(shape, coef, u, SOLAR_E, 0);
  }
  for (i = 0; i < n24; i++) {
    const real_t d = fabs(u[i] - p->U1[i]);
    if (d > maxe) maxe = d;
    diff_l1 += d;
    if (fabs((double)u[i]) > mref) mref = fabs((double)u[i]);
  }
  *max_err = maxe; *l1_err = diff_l1; *max_ref = mref;
  free(u);
  free(coef);
  gpu_detach(p);
  arrays_free(p);
  return (diff_l1 > 0.0) || (diff_l1 * 0 != 0) || (diff_l1 != diff_l1);
}
