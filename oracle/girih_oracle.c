/*
 * girih_oracle.c -- CPU restatement of the GIRIH star-stencil time stepper.
 *
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this library; nothing under
 * girih_b200/ links, imports or executes it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this file
 * against (a) raw U1 dumps produced by the unmodified reference time steppers
 * (oracle/_ref/ref_dump_*, built from /root/reference by oracle/Makefile) and
 * (b) the committed fixtures under tests/golden/ that were generated the same
 * way (tests/golden/make_golden.py).  The reference itself holds no golden
 * vectors; its own criterion is "optimised stepper ==
 * serial reference kernels, bit for bit" (src/verification.c:842), which the
 * reference binaries built here satisfy (eMax 0) for kernels 0-5 and 7.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  Arithmetic is IEEE add/mul in the reference's evaluation
 * order; the file must be compiled with -ffp-contract=off (oracle/Makefile).
 * Loops are OpenMP-parallel over k only: each point is one expression on the
 * same inputs, so threading cannot change a bit.
 *
 * Compiled twice (ORACLE_DP = 0 / 1); exported names carry an _sp / _dp suffix.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifndef ORACLE_DP
#define ORACLE_DP 1
#endif

/* ORACLE_FMA=1: the same source compiled with -mfma -ffp-contract=fast (suffix _dpf/_spf): gcc then
 * contracts the expressions below exactly as it contracts the reference's FUNC_BODY in the *_fast
 * reference builds (oracle/Makefile RFAST) -- the checker for the library's "contract" option. */
#ifndef ORACLE_FMA
#define ORACLE_FMA 0
#endif

#if ORACLE_DP
typedef double real_t;                 /* src/data_structures.h:99-101 */
#if ORACLE_FMA
#define SFX(name) name##_dpf
#else
#define SFX(name) name##_dp
#endif
#else
typedef float real_t;                  /* src/data_structures.h:102-105 */
#if ORACLE_FMA
#define SFX(name) name##_spf
#else
#define SFX(name) name##_sp
#endif
#endif

/* operator registry, src/kernels/stencils.c:260-271 */
enum { COEF_CONST = 0, COEF_VAR = 1, COEF_AXSYM = 2, COEF_NOSYM = 3, COEF_SOLAR = 4 };
typedef struct { int r, time_order, nd, coeff, is_box; } kinfo_t;
static const kinfo_t KINFO[8] = {
  {4, 2, 3,  COEF_CONST, 0},   /* 0 iso_ref                          */
  {1, 1, 2,  COEF_CONST, 0},   /* 1 iso_ref_2space_1time             */
  {1, 1, 4,  COEF_VAR,   0},   /* 2 iso_ref_2space_1time_var         */
  {1, 1, 6,  COEF_AXSYM, 0},   /* 3 iso_ref_2space_1time_var_axsym   */
  {4, 1, 15, COEF_AXSYM, 0},   /* 4 iso_ref_8space_1time_var_axsym   */
  {1, 1, 9,  COEF_NOSYM, 0},   /* 5 iso_ref_2space_1time_var_nosym   */
  {1, 1, 40, COEF_SOLAR, 0},   /* 6 solar: oracle_solar_* below       */
  {1, 1, 2,  COEF_CONST, 1},   /* 7 box_ref_2space_1time             */
};

/* default coefficient list, src/utils.c:116-119 (double literals cast to real_t) */
static const double G_COEF[11] = {-0.28472, 0.16000, -0.02000, 0.00254,
    -0.00018, -0.18472, 0.19, -0.0500, 0.00554, -0.0009, 0.00354};

#if ORACLE_DP && !ORACLE_FMA
/* kernel_info(k, out[5]) -> r, time_order, nd, coeff kind, is_box; returns 0 if k valid */
int oracle_kernel_info(int k, int out[5])
{
  if (k < 0 || k > 7) return -1;
  out[0] = KINFO[k].r; out[1] = KINFO[k].time_order; out[2] = KINFO[k].nd;
  out[3] = KINFO[k].coeff; out[4] = KINFO[k].is_box;
  return 0;
}

/* domain shape of one (sub)domain, src/utils.c:367-374.  `alignment` is in ELEMENTS
 * (the reference applies the same number as bytes to posix_memalign and as elements
 * to the padding computation). */
void oracle_domain_shape(const int lstencil[3], int r, int alignment, int padding, int out[3])
{
  int pad = 0;
  if (padding) {
    int comp = (lstencil[0] + 2 * r) % alignment;
    if (comp != 0) pad = alignment - comp;
  }
  out[0] = lstencil[0] + 2 * r + pad;
  out[1] = lstencil[1] + 2 * r;
  out[2] = lstencil[2] + 2 * r;
}

/* 1-D block decomposition with remainder cells on the low ranks, src/utils.c:339-356 */
void oracle_decompose(int n, int nparts, int coord, int *local_n, int *gb)
{
  if (nparts > 1) {
    int q = n / nparts, rem = n % nparts;
    if (coord < rem) { *local_n = q + 1; *gb = coord * (q + 1); }
    else             { *local_n = q;     *gb = rem * (q + 1) + (coord - rem) * q; }
  } else { *local_n = n; *gb = 0; }
}

/* number of real_t in the coefficient array, src/utils.c:182-207 */
uint64_t oracle_coef_size(int k, uint64_t ln_domain)
{
  int r = KINFO[k].r;
  switch (KINFO[k].coeff) {
    case COEF_CONST: return 10;
    case COEF_VAR:   return ln_domain * (uint64_t)(1 + r);
    case COEF_AXSYM: return ln_domain * (uint64_t)(1 + 3 * r);
    case COEF_NOSYM: return ln_domain * (uint64_t)(1 + 6 * r);
    case COEF_SOLAR: return ln_domain * 28u * 2u;                       /* utils.c:199-201 */
    default:         return 0;
  }
}

/* the diamond stepper's nt rounding, src/kernels/diamond_utils.c:1042-1056 (REGULAR type) */
int oracle_diamond_round_nt(int nt, int t_dim)
{
  int remain = (nt - 2) % ((t_dim + 1) * 2);
  if (remain != 0) nt = nt + (t_dim + 1) * 2 - remain;
  return nt;
}
#endif /* ORACLE_DP (precision-independent helpers are emitted once) */

/* coefficient fill, src/utils.c:436-481 */
void SFX(oracle_init_coeff)(int k, uint64_t ln_domain, real_t *coef)
{
  uint64_t i, m, ax;
  int r = KINFO[k].r;
  real_t g[11];
  for (i = 0; i < 11; i++) g[i] = (real_t)G_COEF[i];          /* utils.c:119 */
  switch (KINFO[k].coeff) {
    case COEF_CONST:                                            /* utils.c:441-444 */
      for (i = 0; i < (uint64_t)r + 1; i++) coef[i] = g[i];
      break;
    case COEF_VAR:                                              /* utils.c:446-452 */
      for (m = 0; m <= (uint64_t)r; m++)
        for (i = 0; i < ln_domain; i++) coef[i + m * ln_domain] = g[m];
      break;
    case COEF_AXSYM:                                            /* utils.c:454-466 */
      for (i = 0; i < ln_domain; i++) coef[i] = g[0];
      for (m = 0; m < (uint64_t)r; m++)
        for (ax = 0; ax < 3; ax++)
          for (i = 0; i < ln_domain; i++)
            coef[i + ln_domain + 3 * m * ln_domain + ax * ln_domain] = g[m + 1];
      break;
    case COEF_NOSYM:                                            /* utils.c:468-481 */
      for (i = 0; i < ln_domain; i++) coef[i] = g[0];
      for (m = 0; m < (uint64_t)r; m++)
        for (ax = 0; ax < 3; ax++)
          for (i = 0; i < ln_domain; i++) {
            coef[i + ln_domain + 6 * m * ln_domain + 2 * ax * ln_domain] = g[m + 1];
            coef[i + ln_domain + 6 * m * ln_domain + (2 * ax + 1) * ln_domain] = g[m + 1];
          }
      break;
    default: break;
  }
}

/*
 * Field fill of one (sub)domain, src/utils.c:605-697 (domain_data_fill_std).
 *   dshape     local domain shape incl. frame/halo and x padding
 *   lstencil   local interior shape, gstencil the GLOBAL interior shape
 *   gb         global begin of this subdomain (utils.c:345-353)
 *   first/last flags per axis: rank_coords[d]==0 / ==shape[d]-1
 * U3 may be NULL (time_order 1).  For the undecomposed domain (gb=0, all flags 1)
 * this equals the verifier's fill, src/verification.c:223-250.
 */
void SFX(oracle_fill)(const int dshape[3], const int lstencil[3], const int gstencil[3],
                      const int gb[3], const int first[3], const int last[3], int r,
                      real_t *U1, real_t *U2, real_t *U3)
{
  const int nnx = dshape[0], nny = dshape[1], nnz = dshape[2];
  const uint64_t n = (uint64_t)nnx * nny * nnz;
  uint64_t i;
  int x, y, z, xb = 0, yb = 0, zb = 0;
  int xe = lstencil[0] + 2 * r, ye = lstencil[1] + 2 * r, ze = lstencil[2] + 2 * r;
  for (i = 0; i < n; i++) { U1[i] = 0.0; U2[i] = 0.0; if (U3) U3[i] = 0.0; }  /* :610-616 */
  if (first[0]) xb += r;                                                        /* :623-625 */
  if (first[1]) yb += r;
  if (first[2]) zb += r;
  if (last[0]) xe -= r;                                                         /* :626-628 */
  if (last[1]) ye -= r;
  if (last[2]) ze -= r;
  for (z = zb; z < ze; z++)
    for (y = yb; y < ye; y++)
      for (x = xb; x < xe; x++) {
        uint64_t gi = (uint64_t)x + gb[0], gj = (uint64_t)y + gb[1], gk = (uint64_t)z + gb[2];
        /* :638 -- the sum is rounded to real_t BEFORE the *1.845703 (real_t r) */
        real_t rr = 1.0 / 3 * (1.0 * gi / gstencil[0] + 1.0 * gj / gstencil[1] + 1.0 * gk / gstencil[2]);
        uint64_t idx = ((uint64_t)z * nny + y) * nnx + x;
        U1[idx] = rr * 1.845703;                                                /* :640-643 */
        U2[idx] = rr * 1.845703;
        if (U3) U3[idx] = rr * 1.845703;
      }
  if (first[0])                                                                 /* :679-687 */
    for (z = 0; z < nnz; z++)
      for (y = 0; y < nny; y++) {
        uint64_t idx = ((uint64_t)z * nny + y) * nnx;
        U1[idx] += 100.1; U2[idx] += 100.1;
      }
  if (last[0])                                                                  /* :688-696 */
    for (z = 0; z < nnz; z++)
      for (y = 0; y < nny; y++) {
        uint64_t idx = ((uint64_t)z * nny + y) * nnx + (lstencil[0] + 2 * r - 1);
        U1[idx] += 100.1; U2[idx] += 100.1;
      }
}

/* ------------------------------------------------------------------------------------------
 * One time step over the box [xb,xe) x [yb,ye) x [zb,ze): the loop nest of
 * src/kernels/stencils_spt_blk.ic:19-48 around each FUNC_BODY of src/kernels/stencils.c.
 * (y blocking / OpenMP scheduling of the template do not affect results.)
 * ------------------------------------------------------------------------------------------ */
#define IDX(i,j,k) (((uint64_t)(k) * nny + (j)) * nnx + (i))
#define V(i,j,k)   v[IDX(i,j,k)]
#define C(m)       coef[(m) * ln + c]

void SFX(oracle_step)(int kern, const int shape[3], int xb, int yb, int zb, int xe, int ye, int ze,
                      const real_t *coef, real_t *u, const real_t *v, const real_t *roc2)
{
  const int nnx = shape[0], nny = shape[1];
  const uint64_t ln = (uint64_t)shape[0] * shape[1] * shape[2];
  int k;
#pragma omp parallel for schedule(static)
  for (k = zb; k < ze; k++) {
    int i, j;
    for (j = yb; j < ye; j++) {
      for (i = xb; i < xe; i++) {
        const uint64_t c = IDX(i, j, k);
        switch (kern) {
          case 0:   /* stencils.c:27-41 */
            u[c] = ((real_t)(2.0)) * v[c] - u[c] + roc2[c] * (coef[0] * v[c]
                 + coef[1] * (V(i+1,j,k) + V(i-1,j,k))
                 + coef[1] * (V(i,j+1,k) + V(i,j-1,k))
                 + coef[1] * (V(i,j,k+1) + V(i,j,k-1))
                 + coef[2] * (V(i+2,j,k) + V(i-2,j,k))
                 + coef[2] * (V(i,j+2,k) + V(i,j-2,k))
                 + coef[2] * (V(i,j,k+2) + V(i,j,k-2))
                 + coef[3] * (V(i+3,j,k) + V(i-3,j,k))
                 + coef[3] * (V(i,j+3,k) + V(i,j-3,k))
                 + coef[3] * (V(i,j,k+3) + V(i,j,k-3))
                 + coef[4] * (V(i+4,j,k) + V(i-4,j,k))
                 + coef[4] * (V(i,j+4,k) + V(i,j-4,k))
                 + coef[4] * (V(i,j,k+4) + V(i,j,k-4)));
            break;
          case 1:   /* stencils.c:73-78 */
            u[c] = coef[0] * v[c]
                 + coef[1] * (V(i+1,j,k) + V(i-1,j,k))
                 + coef[1] * (V(i,j+1,k) + V(i,j-1,k))
                 + coef[1] * (V(i,j,k-1) + V(i,j,k+1));
            break;
          case 2:   /* stencils.c:94-99 */
            u[c] = C(0) * v[c]
                 + C(1) * (V(i+1,j,k) + V(i-1,j,k))
                 + C(1) * (V(i,j+1,k) + V(i,j-1,k))
                 + C(1) * (V(i,j,k+1) + V(i,j,k-1));
            break;
          case 3:   /* stencils.c:121-126 */
            u[c] = C(0) * v[c]
                 + C(1) * (V(i+1,j,k) + V(i-1,j,k))
                 + C(2) * (V(i,j+1,k) + V(i,j-1,k))
                 + C(3) * (V(i,j,k+1) + V(i,j,k-1));
            break;
          case 4:   /* stencils.c:148-162 */
            u[c] = C(0)  * v[c]
                 + C(1)  * (V(i+1,j,k) + V(i-1,j,k))
                 + C(2)  * (V(i,j+1,k) + V(i,j-1,k))
                 + C(3)  * (V(i,j,k+1) + V(i,j,k-1))
                 + C(4)  * (V(i+2,j,k) + V(i-2,j,k))
                 + C(5)  * (V(i,j+2,k) + V(i,j-2,k))
                 + C(6)  * (V(i,j,k+2) + V(i,j,k-2))
                 + C(7)  * (V(i+3,j,k) + V(i-3,j,k))
                 + C(8)  * (V(i,j+3,k) + V(i,j-3,k))
                 + C(9)  * (V(i,j,k+3) + V(i,j,k-3))
                 + C(10) * (V(i+4,j,k) + V(i-4,j,k))
                 + C(11) * (V(i,j+4,k) + V(i,j-4,k))
                 + C(12) * (V(i,j,k+4) + V(i,j,k-4));
            break;
          case 5:   /* stencils.c:193-201 */
            u[c] = C(0) * v[c]
                 + C(1) * V(i-1,j,k)
                 + C(2) * V(i+1,j,k)
                 + C(3) * V(i,j-1,k)
                 + C(4) * V(i,j+1,k)
                 + C(5) * V(i,j,k-1)
                 + C(6) * V(i,j,k+1);
            break;
          case 7:   /* stencils.c:227-242 */
            u[c] = coef[0] * v[c]
                 + coef[1] * (V(i+1,j,k) + V(i-1,j,k))
                 + coef[1] * (V(i,j+1,k) + V(i,j-1,k))
                 + coef[1] * (V(i,j,k-1) + V(i,j,k+1))
                 + coef[2] * (V(i+1,j,k-1) + V(i-1,j,k-1))
                 + coef[2] * (V(i,j+1,k-1) + V(i,j-1,k-1))
                 + coef[2] * (V(i+1,j+1,k) + V(i-1,j-1,k))
                 + coef[2] * (V(i+1,j-1,k) + V(i-1,j+1,k))
                 + coef[2] * (V(i+1,j,k+1) + V(i-1,j,k+1))
                 + coef[2] * (V(i,j+1,k+1) + V(i,j-1,k+1))
                 + coef[3] * (V(i+1,j+1,k+1) + V(i-1,j-1,k-1))
                 + coef[3] * (V(i+1,j-1,k+1) + V(i-1,j+1,k-1))
                 + coef[3] * (V(i-1,j-1,k+1) + V(i+1,j+1,k-1))
                 + coef[3] * (V(i-1,j+1,k+1) + V(i+1,j-1,k-1));
            break;
          default: break;
        }
      }
    }
  }
}

/*
 * The reference time loop on one undecomposed domain: src/kernels/nb_naive_ts.c:187-203
 * (== the verifier's loop, src/verification.c:281-284):
 *     for (it = 0; it < nt; it += 2) { U1 <- step(U2); U2 <- step(U1); }
 * so an odd nt executes nt+1 steps, and afterwards U1 = level nt-1 (nt even).
 * The box is xb=yb=zb=r, xe=nx+r, ye=nny-r, ze=nnz-r (nb_naive_ts.c:189).
 */
void SFX(oracle_run_naive)(int kern, const int shape[3], int nx, int nt,
                           const real_t *coef, real_t *U1, real_t *U2, const real_t *U3)
{
  const int r = KINFO[kern].r;
  int it;
  for (it = 0; it < nt; it += 2) {
    SFX(oracle_step)(kern, shape, r, r, r, nx + r, shape[1] - r, shape[2] - r, coef, U1, U2, U3);
    SFX(oracle_step)(kern, shape, r, r, r, nx + r, shape[1] - r, shape[2] - r, coef, U2, U1, U3);
  }
}

/*
 * Exactly `nsteps` steps with the same parity convention (odd global steps write U1):
 * what the diamond stepper leaves behind after its nt-1 executed steps
 * (src/kernels/diamond_ts.c:440-444, 871-946): U1 = level nt-1, U2 = level nt-2.
 */
void SFX(oracle_run_steps)(int kern, const int shape[3], int nx, int nsteps,
                           const real_t *coef, real_t *U1, real_t *U2, const real_t *U3)
{
  const int r = KINFO[kern].r;
  int s;
  for (s = 1; s <= nsteps; s++) {
    if (s % 2 == 1)
      SFX(oracle_step)(kern, shape, r, r, r, nx + r, shape[1] - r, shape[2] - r, coef, U1, U2, U3);
    else
      SFX(oracle_step)(kern, shape, r, r, r, nx + r, shape[1] - r, shape[2] - r, coef, U2, U1, U3);
  }
}

/*
 * Comparator, src/verification.c:823-860 (compare_results_std): max |ref-target| and the
 * L1 sum over the interior, accumulated in real_t like the reference; additionally the
 * max |ref| so callers can form the relative L-infinity the north star states.
 * `target` is the gathered interior (nx*ny*nz, no halo), `ref` the full reference domain.
 * Returns 0 when the reference's PASS criterion holds (diff_l1 == 0 and finite).
 */
int SFX(oracle_compare)(const real_t *ref, const real_t *target, int nx, int ny, int nz, int r,
                        double *max_err, double *l1_err, double *max_ref)
{
  const int nnx = nx + 2 * r, nny = ny + 2 * r;
  real_t diff_l1 = 0.0, maxe = 0.0;
  double mref = 0.0;
  int i, j, k;
  for (k = 0; k < nz; k++)
    for (j = 0; j < ny; j++)
      for (i = 0; i < nx; i++) {
        real_t a = ref[IDX(i + r, j + r, k + r)];
        real_t d = fabs(a - target[((uint64_t)k * ny + j) * nx + i]);
        if (d > maxe) maxe = d;
        diff_l1 += d;
        if (fabs((double)a) > mref) mref = fabs((double)a);
      }
  *max_err = maxe; *l1_err = diff_l1; *max_ref = mref;
  return ((diff_l1 > 0.0) || (diff_l1 * 0 != 0) || (diff_l1 != diff_l1)) ? 1 : 0;
}


/* ------------------------------------------------------------------------------------------
 * Table slot 6, "solar": 12 complex field components (6 H, 6 E) in ONE array u of
 * 12 x nnx*nny*nnz complex numbers (re, im interleaved; src/utils.c:168-172) and 28 complex
 * coefficient arrays (src/utils.c:199-201).  One time step = the H update of every interior
 * cell followed by the E update of every interior cell, in place
 * (src/kernels/solar_spt_blk.ic:20-200 / :202-386 / :388-397; the serial verifier
 * src/verification.c:481-784 holds the same expressions, character for character).
 * Pinned like the rest: tests/test_oracle_vs_reference.py against oracle/_ref dumps and the
 * committed fixtures.
 * ------------------------------------------------------------------------------------------ */
/* coefficients, src/utils.c:483-489: the flat index modulo 10 picks the value */
void SFX(oracle_solar_init_coeff)(uint64_t ln_domain, real_t *coef)
{
  uint64_t i, n = ln_domain * 28u * 2u;
  for (i = 0; i < n; i++) coef[i] = (real_t)G_COEF[i % 10];
}

/* fields, src/utils.c:698-727 (== the verifier's fill, src/verification.c:253-268, for gb = 0):
 * every cell of the local array incl. its frame, no zero padding, no source planes */
void SFX(oracle_solar_fill)(const int dshape[3], const int gstencil[3], const int gb[3], real_t *u)
{
  const int nnx = dshape[0], nny = dshape[1], nnz = dshape[2];
  const uint64_t n = (uint64_t)nnx * nny * nnz;
  int f, x, y, z;
  for (f = 0; f < 12; f++)
    for (z = 0; z < nnz; z++)
      for (y = 0; y < nny; y++)
        for (x = 0; x < nnx; x++) {
          uint64_t gi = (uint64_t)x + gb[0], gj = (uint64_t)y + gb[1], gk = (uint64_t)z + gb[2];
          real_t rr = 1.0 / (3.0) * (1.0 * gi / gstencil[0] + 1.0 * gj / gstencil[1] + 1.0 * gk / gstencil[2]);
          uint64_t idx = 2 * ((((uint64_t)z * nny + y) * nnx + x) + n * f);
          u[idx] = rr * 1.845703;
          u[idx + 1] = rr * 1.845703 / 3.0;
        }
}

/* the four orders in which the reference writes the staggered difference of two source fields */
#define SD_A(P, Q, o) (P[i + (o)] - P[s + (o)] + Q[i + (o)] - Q[s + (o)])   /* cur - sub + cur - sub */
#define SD_B(P, Q, o) (P[s + (o)] - P[i + (o)] + Q[s + (o)] - Q[i + (o)])   /* sub - cur + sub - cur */
#define SD_C(P, Q, o) (P[s + (o)] + Q[s + (o)] - P[i + (o)] - Q[i + (o)])   /* sub + sub - cur - cur */
#define SD_D(P, Q, o) (P[i + (o)] + Q[i + (o)] - P[s + (o)] - Q[s + (o)])   /* cur + cur - sub - sub */
#define FLD(m) (u + (uint64_t)(m) * ln2)
#define COE(m) (coef + (uint64_t)(m) * ln2)
/* H component F (coefficients c = F, t = 6 + F), sources P, Q at offset OFF cells, difference SD */
#define H_UPD(F, P, Q, OFF, SD, BND) do {                                                   \
    real_t *h = FLD(F); const real_t *pp = FLD(P), *qq = FLD(Q);                             \
    const real_t *cc = COE(F), *tt = COE(6 + (F));                                           \
    const uint64_t s = i + 2 * (OFF);                                                        \
    const real_t dR = SD(pp, qq, 0), dI = SD(pp, qq, 1);                                     \
    real_t asgn;                                                                             \
    if ((BND) >= 0) {                                                                        \
      const real_t *bb = COE((BND) >= 0 ? (BND) : 0);                                        \
      asgn     = h[i] * tt[i] - h[i + 1] * tt[i + 1] + bb[i] - cc[i] * dR + cc[i + 1] * dI;  \
      h[i + 1] = h[i] * tt[i + 1] + h[i + 1] * tt[i] + bb[i + 1] - cc[i] * dI - cc[i + 1] * dR; \
    } else {                                                                                 \
      asgn     = h[i] * tt[i] - h[i + 1] * tt[i + 1] - cc[i] * dR + cc[i + 1] * dI;          \
      h[i + 1] = h[i] * tt[i + 1] + h[i + 1] * tt[i] - cc[i] * dI - cc[i + 1] * dR;          \
    }                                                                                        \
    h[i] = asgn;                                                                             \
  } while (0)
/* E component F (6..11; coefficients c = 14 + F - 6, t = 20 + F - 6) */
#define E_UPD(F, P, Q, OFF, SD, BND) do {                                                   \
    real_t *e = FLD(F); const real_t *pp = FLD(P), *qq = FLD(Q);                             \
    const real_t *cc = COE(14 + (F) - 6), *tt = COE(20 + (F) - 6);                           \
    const uint64_t s = i + 2 * (OFF);                                                        \
    const real_t dR = SD(pp, qq, 0), dI = SD(pp, qq, 1);                                     \
    real_t asgn;                                                                             \
    if ((BND) >= 0) {                                                                        \
      const real_t *bb = COE((BND) >= 0 ? (BND) : 0);                                        \
      asgn     = e[i] * tt[i] - e[i + 1] * tt[i + 1] + bb[i] + cc[i] * dR - cc[i + 1] * dI;  \
      e[i + 1] = e[i] * tt[i + 1] + e[i + 1] * tt[i] + bb[i + 1] + cc[i] * dI + cc[i + 1] * dR; \
    } else {                                                                                 \
      asgn     = e[i] * tt[i] - e[i + 1] * tt[i + 1] + cc[i] * dR - cc[i + 1] * dI;          \
      e[i + 1] = e[i] * tt[i + 1] + e[i + 1] * tt[i] + cc[i] * dI + cc[i + 1] * dR;          \
    }                                                                                        \
    e[i] = asgn;                                                                             \
  } while (0)

/* field numbers, src/kernels/solar_spt_blk.ic:29-41; coefficient numbers :44-60, :221-236 */
enum { S_HYX = 0, S_HZX, S_HXY, S_HZY, S_HXZ, S_HYZ, S_EXZ, S_EYZ, S_EYX, S_EZX, S_EXY, S_EZY };
enum { S_HXBND = 12, S_HYBND = 13, S_EXBND = 26, S_EYBND = 27 };

/* which = 1: H update, 2: E update, 3: both (ALL_FIELDS), over the box [xb,xe) x [yb,ye) x [zb,ze) */
void SFX(oracle_solar_step)(const int shape[3], int xb, int yb, int zb, int xe, int ye, int ze,
                            const real_t *coef, real_t *u, int which)
{
  const int nnx = shape[0], nny = shape[1];
  const uint64_t ln2 = 2 * (uint64_t)shape[0] * shape[1] * shape[2];
  const int64_t SX = 1, SY = nnx, SZ = (int64_t)nnx * nny;
  int x, j, k;
  if (which & 1) {
#pragma omp parallel for private(x, j) schedule(static)
    for (k = zb; k < ze; k++)
      for (j = yb; j < ye; j++)
        for (x = xb; x < xe; x++) {
          const uint64_t i = 2 * (((uint64_t)k * nny + j) * nnx + x);
          H_UPD(S_HYX, S_EXY, S_EXZ, -SZ, SD_A, S_HYBND);   /* solar_spt_blk.ic:78-84 */
          H_UPD(S_HZX, S_EXY, S_EXZ, -SY, SD_B, -1);        /* :100-106 */
          H_UPD(S_HXY, S_EYX, S_EYZ, -SZ, SD_B, S_HXBND);   /* :122-128 */
          H_UPD(S_HZY, S_EYX, S_EYZ, -SX, SD_A, -1);        /* :144-150 */
          H_UPD(S_HXZ, S_EZX, S_EZY, -SY, SD_A, -1);        /* :166-172 */
          H_UPD(S_HYZ, S_EZX, S_EZY, -SX, SD_C, -1);        /* :188-194 */
        }
  }
  if (which & 2) {
#pragma omp parallel for private(x, j) schedule(static)
    for (k = zb; k < ze; k++)
      for (j = yb; j < ye; j++)
        for (x = xb; x < xe; x++) {
          const uint64_t i = 2 * (((uint64_t)k * nny + j) * nnx + x);
          E_UPD(S_EXZ, S_HZX, S_HZY, +SY, SD_B, -1);        /* solar_spt_blk.ic:263-269 */
          E_UPD(S_EYZ, S_HZX, S_HZY, +SX, SD_D, -1);        /* :285-291 */
          E_UPD(S_EYX, S_HXY, S_HXZ, +SZ, SD_B, S_EYBND);   /* :307-313 */
          E_UPD(S_EZX, S_HXY, S_HXZ, +SY, SD_D, -1);        /* :329-335 */
          E_UPD(S_EXY, S_HYX, S_HYZ, +SZ, SD_A, S_EXBND);   /* :351-357 */
          E_UPD(S_EZY, S_HYX, S_HYZ, +SX, SD_B, -1);        /* :373-379 */
        }
  }
}

/* `nsteps` time steps on the undecomposed domain: what nb_naive_ts.c:187-203 does with U2 == 0
 * (solar() falls back to the one array it has, solar_spt_blk.ic:392) -- nt calls, nt rounded up to even
 * by the caller like oracle_run_naive */
void SFX(oracle_solar_run)(const int shape[3], int nx, int nsteps, const real_t *coef, real_t *u)
{
  int s;
  for (s = 0; s < nsteps; s++)
    SFX(oracle_solar_step)(shape, 1, 1, 1, nx + 1, shape[1] - 1, shape[2] - 1, coef, u, 3);
}
