/*
 * init.c -- decomposition, shapes, allocation, coefficient and field initialisation, and the
 * attach/detach of the device context.
 * Follows src/utils.c:153-245 (arrays_allocate/free), :253-300 (set_kernels), :317-433 (init),
 * :436-496 (init_coeff), :605-697 (domain_data_fill_std) and the validity rules of
 * src/kernels/diamond_utils.c:850-1099 (intra_diamond_info_init) that affect results.
 */
#define _POSIX_C_SOURCE 200112L
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "girih_host.h"

/* The reference's RAISE_ERROR (src/data_structures.h:299-314): message on stderr, exit(1). */
void girih_fatal(const Parameters *p, const char *fmt, ...) {
  /* ranks are host threads of one process (team.c): whichever rank gets here first reports, once, before any
   * rank can end the process -- waiting for rank 0 would race with another rank's exit() */
  static int reported = 0;
  (void)p;
  if (__sync_lock_test_and_set(&reported, 1) == 0) {
    va_list ap;
    fprintf(stderr, "ERROR: ");
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\n");
    fflush(stderr);
    exit(1);
  }
  for (;;) pause(); /* another rank is reporting: let it finish the message and end the process */
}

static void set_kernels(Parameters *p) { /* src/utils.c:253-262 */
  girih_kernel_desc d;
  if (girih_kernel_info(p->target_kernel, &d) != GIRIH_OK) girih_fatal(p, "unknown stencil kernel %d", p->target_kernel);
  if (!d.gpu_supported) { /* src/kernels/stencils.h:40-47 */
    if (p->mpi_rank == 0) printf("ERROR: unsupported configuration for the selected stencil\n");
    exit(1);
  }
  p->stencil.name = d.name;
  p->stencil.r = d.r;
  p->stencil.time_order = d.time_order;
  p->stencil.nd = d.nd;
  p->stencil.shape = d.shape;
  p->stencil.coeff = d.coeff;
  if (d.coeff == GIRIH_COEF_SOLAR) {
    /* the reference's default wavefront has no solar kernel (mwd_func_list, src/kernels/stencils.c:303-311:
     * not_supported_mwd, src/kernels/stencils.h:40-43) and its halo exchange moves one real per cell */
    if (p->target_ts == 2) {
      if (p->mpi_rank == 0) printf("ERROR: unsupported configuration for the selected stencil\n");
      exit(1);
    }
    if (p->mpi_size > 1) girih_fatal(p, "the solar kernel runs on one GPU (--npx 1 --npy 1 --npz 1)");
  }
}

/* src/kernels/diamond_utils.c:850-1099, the parts that are not CPU thread/cache tuning */
static void intra_diamond_info_init(Parameters *p) {
  const int r = p->stencil.r;
  int remain, nt = p->nt, diam_width, diam_concurrency;

  if ((p->mwd_type > 1) && (r > 1)) /* :859 */
    girih_fatal(p, "Relaxed synchronization implementations are disabled for stencil radius > 1 for performance reasons");
  if (p->mwd_type < 0 || p->mwd_type > 3) girih_fatal(p, "unknown MWD implementation %d", p->mwd_type);

  if (p->t_dim == -1) {
    /* the reference auto-tunes the time unroll here (:880-888); the GPU sweep does not depend on
     * it, so take the largest diamond that tiles the domain */
    int cand;
    for (cand = 7; cand >= 1; cand -= 2)
      if (p->lstencil_shape[1] % ((cand + 1) * 2 * r) == 0 && p->stencil_shape[2] >= cand * 2 * r + 1) break;
    p->t_dim = cand < 1 ? 1 : cand;
    if (p->mpi_rank == 0 && p->verbose == 1) printf("[AUTO TUNE] selected time unroll (t_dim): %d\n", p->t_dim);
  }
  if (p->thread_group_size == -1) p->thread_group_size = 1;
  if (p->num_wf == -1) p->num_wf = 1;

  if (p->t_dim < 1) girih_fatal(p, "Diamond method does not support unrolling in time less than 1"); /* :1023 */
  if (p->t_dim % 2 == 0) girih_fatal(p, "diamond method does not supports even time unrolling");     /* :1029 */
  if (p->t.shape[0] > 1 || p->t.shape[1] > 1) /* the reference's rule is npx == npz == 1 (:1035-1040); the GPU
                                                * diamond stepper shards z-slabs instead of y-diamonds */
    girih_fatal(p, "the Diamond stepper of this build decomposes the domain across the Z direction only (use --npx 1 --npy 1 --npz <GPUs>)");

  diam_width = (p->t_dim + 1) * 2 * r;
  diam_concurrency = p->lstencil_shape[1] / diam_width;
  if (p->wavefront == 1) { /* :952-960 */
    const int min_z = (p->t_dim * 2) * r + 1;
    if (p->stencil_shape[2] < min_z)
      girih_fatal(p, "The single core wavefront requires a minimum size of %d at the Z direction in the current configurations", min_z);
  }
  /* :1011-1012 */
  p->idiamond_pro_epi_logue_updates = 1ULL * p->stencil_shape[0] * p->stencil_shape[2] * 2ULL * diam_concurrency *
                                      ((p->t_dim + 1) * (p->t_dim + 1) + (p->t_dim + 1)) * r;
  /* round the number of time steps to the nearest valid number, :1042-1056 */
  remain = (nt - 2) % ((p->t_dim + 1) * 2);
  if (remain != 0) {
    const int nt2 = nt + (p->t_dim + 1) * 2 - remain;
    if (nt2 != nt) {
      if (p->mpi_rank == 0 && p->verbose == 1)
        printf("###INFO: Modified nt from %03d to %03d for the intra-diamond method to work properly\n", nt, nt2);
      p->nt = nt2;
    }
  }
  if (p->lstencil_shape[1] < diam_width) /* :1085 */
    girih_fatal(p, "Intra-diamond method requires the sub-domain size to fit at least one diamond: %d elements in Y [stencil_radius*2*(time_unrolls+1)]. Given %d elements",
                diam_width, p->lstencil_shape[1]);
  if (p->lstencil_shape[1] % diam_width != 0) /* :1092 */
    girih_fatal(p, "Intra-diamond method requires the sub-domain size to be multiples of the diamond width: %d elements [stencil_radius*2*(time_unrolls+1)]",
                diam_width);
}

void init(Parameters *p) { /* src/utils.c:317-433 */
  int i, q, rem, padding_comp, padding_size = 0;

  if (p->target_ts < 0 || p->target_ts > 2) girih_fatal(p, "unknown time stepper %d", p->target_ts);
  set_kernels(p);
  p->n_stencils = (uint64_t)p->stencil_shape[0] * p->stencil_shape[1] * p->stencil_shape[2];
  if (p->mpi_size == 1 && p->halo_concat == 1) p->halo_concat = 0;

  /* MPI_Cart_create / MPI_Cart_coords (src/mpi_utils.c:74-77): row-major rank order, z fastest */
  p->t.rank_coords[2] = p->mpi_rank % p->t.shape[2];
  p->t.rank_coords[1] = (p->mpi_rank / p->t.shape[2]) % p->t.shape[1];
  p->t.rank_coords[0] = p->mpi_rank / (p->t.shape[2] * p->t.shape[1]);
  for (i = 0; i < 3; i++) { /* :339-356 */
    if (p->t.shape[i] > 1) {
      q = p->stencil_shape[i] / p->t.shape[i];
      rem = p->stencil_shape[i] % p->t.shape[i];
      if (p->t.rank_coords[i] < rem) {
        p->lstencil_shape[i] = q + 1;
        p->gb[i] = p->t.rank_coords[i] * (q + 1);
      } else {
        p->lstencil_shape[i] = q;
        p->gb[i] = rem * (q + 1) + (p->t.rank_coords[i] - rem) * q;
      }
    } else {
      p->lstencil_shape[i] = p->stencil_shape[i];
      p->gb[i] = 0;
    }
    p->ge[i] = p->gb[i] + p->lstencil_shape[i] - 1;
  }
  if (p->lstencil_shape[0] < 1 || p->lstencil_shape[1] < 1 || p->lstencil_shape[2] < 1)
    girih_fatal(p, "more GPUs than grid points along one direction");

  if (is_solar(p) && p->array_padding == 1) { /* :359-361 */
    if (p->mpi_rank == 0) fprintf(stdout, "WARNING: solar kernels do not support array padding\n");
    p->array_padding = 0;
  }
  if (p->array_padding == 1) { /* :367-374: alignment counts ELEMENTS here */
    if (p->alignment < 1) girih_fatal(p, "alignment must be positive");
    padding_comp = (p->lstencil_shape[0] + 2 * p->stencil.r) % p->alignment;
    if (padding_comp != 0) padding_size = p->alignment - padding_comp;
  }
  p->ldomain_shape[0] = p->lstencil_shape[0] + 2 * p->stencil.r + padding_size;
  p->ldomain_shape[1] = p->lstencil_shape[1] + 2 * p->stencil.r;
  p->ldomain_shape[2] = p->lstencil_shape[2] + 2 * p->stencil.r;

  if (p->target_ts == 2) intra_diamond_info_init(p);

  p->ln_domain = (uint64_t)p->ldomain_shape[0] * p->ldomain_shape[1] * p->ldomain_shape[2];
  p->ln_stencils = (uint64_t)p->lstencil_shape[0] * p->lstencil_shape[1] * p->lstencil_shape[2];

  if (p->debug == 1) { /* :404-432 */
    int j;
    for (j = 0; j < p->mpi_size; j++) {
      if (j == p->mpi_rank) {
        printf("[%02d]:top(%02d,%02d,%02d)\n", p->mpi_rank, p->t.rank_coords[0], p->t.rank_coords[1], p->t.rank_coords[2]);
        printf("ln_domain:%06llu  lnstencil:%06llu\n", (unsigned long long)p->ln_domain, (unsigned long long)p->ln_stencils);
        printf("  Local stencil Shape:(%03d,%03d,%03d)\n", p->lstencil_shape[0], p->lstencil_shape[1], p->lstencil_shape[2]);
        printf("  Local domain Shape: (%03d,%03d,%03d)\n", p->ldomain_shape[0], p->ldomain_shape[1], p->ldomain_shape[2]);
        printf("  Local begin:        (%03d,%03d,%03d)\n", p->gb[0], p->gb[1], p->gb[2]);
        printf("  Local end:          (%03d,%03d,%03d)\n\n", p->ge[0], p->ge[1], p->ge[2]);
        fflush(stdout);
      }
      team_barrier();
    }
  }
}

uint64_t coef_array_size(const Parameters *p) { /* src/utils.c:182-207 */
  switch (p->stencil.coeff) {
    case GIRIH_COEF_CONSTANT: return 10;
    case GIRIH_COEF_VARIABLE: return p->ln_domain * (uint64_t)(1 + p->stencil.r);
    case GIRIH_COEF_VARIABLE_AXSYM: return p->ln_domain * (uint64_t)(1 + 3 * p->stencil.r);
    case GIRIH_COEF_VARIABLE_NOSYM: return p->ln_domain * (uint64_t)(1 + 6 * p->stencil.r);
    case GIRIH_COEF_SOLAR: return p->ln_domain * 28u * 2u;
    default: return 0;
  }
}

static void *aligned_or_die(const Parameters *p, size_t bytes) {
  void *ptr = NULL;
  /* the reference passes --alignment straight to posix_memalign (src/utils.c:159); round it up to
   * what posix_memalign accepts so that e.g. --alignment 1 still allocates */
  size_t al = (size_t)p->alignment;
  if (al < sizeof(void *)) al = sizeof(void *);
  while (al & (al - 1)) al++;
  if (posix_memalign(&ptr, al, bytes ? bytes : 1) != 0) girih_fatal(NULL, "no sufficient memory");
  return ptr;
}

void arrays_allocate(Parameters *p) { /* src/utils.c:153-218 */
  const uint64_t coef_size = coef_array_size(p);
  if (is_solar(p)) { /* :168-172: one array of 12 complex fields, U2 = 0 */
    const uint64_t domain_size = p->ln_domain * 12u * 2u;
    p->U1 = (real_t *)aligned_or_die(p, sizeof(real_t) * domain_size);
    p->U2 = p->U3 = NULL;
    p->coef = (real_t *)aligned_or_die(p, sizeof(real_t) * coef_size);
    if (p->verbose == 1 && p->mpi_rank == 0)
      printf("[rank=%d] alloc. dom(err=%d):%fGiB coef(err=%d):%fGiB total:%fGiB\n", p->mpi_rank, 0,
             sizeof(real_t) * 1.0 * domain_size / (1024 * 1024 * 1024), 0, sizeof(real_t) * 1.0 * coef_size / (1024 * 1024 * 1024),
             sizeof(real_t) * 1.0 * (coef_size + domain_size) / (1024 * 1024 * 1024));
    return;
  }
  p->U1 = (real_t *)aligned_or_die(p, sizeof(real_t) * p->ln_domain);
  p->U2 = (real_t *)aligned_or_die(p, sizeof(real_t) * p->ln_domain);
  p->U3 = NULL;
  if (p->stencil.time_order == 2) p->U3 = (real_t *)aligned_or_die(p, sizeof(real_t) * p->ln_domain);
  p->coef = (real_t *)aligned_or_die(p, sizeof(real_t) * coef_size);
  /* The reference sets only coef[0..r] of a constant-coefficient operator (src/utils.c:441-444), so
   * the box kernel (slot 7) reads coef[2] and coef[3] from untouched, i.e. zero, fresh pages.  Make
   * that explicit instead of depending on the allocator. */
  if (p->stencil.coeff == GIRIH_COEF_CONSTANT) memset(p->coef, 0, sizeof(real_t) * coef_size);
  if (p->verbose == 1 && (p->mpi_rank == 0 || p->mpi_rank == p->mpi_size - 1))
    printf("[rank=%d] alloc. dom(err=%d):%fGiB coef(err=%d):%fGiB total:%fGiB\n", p->mpi_rank, 0,
           sizeof(real_t) * 2.0 * p->ln_domain / (1024 * 1024 * 1024), 0,
           sizeof(real_t) * 1.0 * coef_size / (1024 * 1024 * 1024),
           sizeof(real_t) * 1.0 * (coef_size + 2 * p->ln_domain) / (1024 * 1024 * 1024));
}

void arrays_free(Parameters *p) { /* src/utils.c:220-245 */
  free(p->coef);
  free(p->U1);
  free(p->U2); /* NULL for the solar slot */
  if (p->stencil.time_order == 2) free(p->U3);
  p->coef = p->U1 = p->U2 = p->U3 = NULL;
}

void init_coeff(Parameters *p) { /* src/utils.c:436-481 */
  uint64_t i, k, ax;
  const uint64_t n = p->ln_domain;
  const int r = p->stencil.r;
  switch (p->stencil.coeff) {
    case GIRIH_COEF_SOLAR: /* :483-489 */
      solar_init_coeff(p);
      break;
    case GIRIH_COEF_CONSTANT:
      for (i = 0; i < (uint64_t)r + 1; i++) p->coef[i] = p->g_coef[i];
      break;
    case GIRIH_COEF_VARIABLE:
      for (k = 0; k <= (uint64_t)r; k++)
        for (i = 0; i < n; i++) p->coef[i + k * n] = p->g_coef[k];
      break;
    case GIRIH_COEF_VARIABLE_AXSYM:
      for (i = 0; i < n; i++) p->coef[i] = p->g_coef[0];
      for (k = 0; k < (uint64_t)r; k++)
        for (ax = 0; ax < 3; ax++)
          for (i = 0; i < n; i++) p->coef[i + n + 3 * k * n + ax * n] = p->g_coef[k + 1];
      break;
    case GIRIH_COEF_VARIABLE_NOSYM:
      for (i = 0; i < n; i++) p->coef[i] = p->g_coef[0];
      for (k = 0; k < (uint64_t)r; k++)
        for (ax = 0; ax < 3; ax++)
          for (i = 0; i < n; i++) {
            p->coef[i + n + 6 * k * n + 2 * ax * n] = p->g_coef[k + 1];
            p->coef[i + n + 6 * k * n + (2 * ax + 1) * n] = p->g_coef[k + 1];
          }
      break;
    default:
      girih_fatal(p, "unknown type of stencil");
  }
}

#define AT(a, i, j, k) ((a)[((uint64_t)(k) * p->ldomain_shape[1] + (j)) * p->ldomain_shape[0] + (i)])

void domain_data_fill(Parameters *p) { /* src/utils.c:605-697 */
  const int r = p->stencil.r;
  uint64_t n;
  int i, j, k, xb = 0, yb = 0, zb = 0;
  int xe = p->lstencil_shape[0] + 2 * r, ye = p->lstencil_shape[1] + 2 * r, ze = p->lstencil_shape[2] + 2 * r;
  if (is_solar(p)) { /* :796-797 */
    solar_domain_fill(p);
    return;
  }
  for (n = 0; n < p->ln_domain; n++) {
    p->U1[n] = 0.0;
    p->U2[n] = 0.0;
    if (p->stencil.time_order == 2) p->U3[n] = 0.0;
  }
  /* the Dirichlet frame stays zero; halos towards a neighbouring rank are filled like interior */
  if (p->t.rank_coords[0] == 0) xb += r;
  if (p->t.rank_coords[1] == 0) yb += r;
  if (p->t.rank_coords[2] == 0) zb += r;
  if (p->t.rank_coords[0] == p->t.shape[0] - 1) xe -= r;
  if (p->t.rank_coords[1] == p->t.shape[1] - 1) ye -= r;
  if (p->t.rank_coords[2] == p->t.shape[2] - 1) ze -= r;
  for (k = zb; k < ze; k++)
    for (j = yb; j < ye; j++)
      for (i = xb; i < xe; i++) {
        const uint64_t gi = (uint64_t)i + p->gb[0], gj = (uint64_t)j + p->gb[1], gk = (uint64_t)k + p->gb[2];
        /* the sum is rounded to real_t before the scaling (real_t r at src/utils.c:607,638) */
        const real_t w = 1.0 / 3 * (1.0 * gi / p->stencil_shape[0] + 1.0 * gj / p->stencil_shape[1] + 1.0 * gk / p->stencil_shape[2]);
        AT(p->U1, i, j, k) = w * 1.845703;
        AT(p->U2, i, j, k) = w * 1.845703;
        if (p->stencil.time_order == 2) AT(p->U3, i, j, k) = w * 1.845703;
      }
  /* source planes at the first and last YZ planes, frame included (src/utils.c:679-696) */
  if (p->t.rank_coords[0] == 0)
    for (k = 0; k < p->ldomain_shape[2]; k++)
      for (j = 0; j < p->ldomain_shape[1]; j++) {
        AT(p->U1, 0, j, k) += BOUNDARY_SRC_VAL;
        AT(p->U2, 0, j, k) += BOUNDARY_SRC_VAL;
      }
  if (p->t.rank_coords[0] == p->t.shape[0] - 1)
    for (k = 0; k < p->ldomain_shape[2]; k++)
      for (j = 0; j < p->ldomain_shape[1]; j++) {
        AT(p->U1, p->lstencil_shape[0] + 2 * r - 1, j, k) += BOUNDARY_SRC_VAL;
        AT(p->U2, p->lstencil_shape[0] + 2 * r - 1, j, k) += BOUNDARY_SRC_VAL;
      }
}

/* ------------------------------------------------------------------------------------------------
 * device context: the only place the host touches include/girih_cuda.h besides the steppers
 * ---------------------------------------------------------------------------------------------- */
static void gpu_check(Parameters *p, int rc, const char *what) {
  if (rc != GIRIH_OK) {
    fprintf(stderr, "ERROR: %s: %s (%s)\n", what, girih_gpu_strerror(rc), p->gpu ? girih_gpu_last_error(p->gpu) : "");
    exit(1);
  }
}

void gpu_attach(Parameters *p) {
  int ndev = 0, rc;
  unsigned char id[GIRIH_COMM_ID_BYTES];
  rc = girih_gpu_count(&ndev);
  if (rc != GIRIH_OK) {
    fprintf(stderr, "ERROR: %s\n", girih_gpu_strerror(rc));
    exit(1);
  }
  if (p->mpi_size > ndev)
    girih_fatal(p, "requested %d GPUs (--npx * --npy * --npz) but only %d are visible", p->mpi_size, ndev);
  p->gpu_device = p->mpi_rank;
  rc = girih_gpu_create(&p->gpu, p->gpu_device, p->target_kernel, (int)sizeof(real_t), p->lstencil_shape,
                        p->ldomain_shape, p->mpi_rank, p->mpi_size);
  gpu_check(p, rc, "girih_gpu_create");
  gpu_check(p, girih_gpu_set_topology(p->gpu, p->t.shape, p->t.rank_coords), "girih_gpu_set_topology");
  if (p->mpi_size > 1) {
    memset(id, 0, sizeof(id));
    if (p->mpi_rank == 0) gpu_check(p, girih_gpu_comm_unique_id(id, sizeof(id)), "girih_gpu_comm_unique_id");
    team_bcast(id, sizeof(id), 0, p->mpi_rank);
    gpu_check(p, girih_gpu_comm_init(p->gpu, id, sizeof(id)), "girih_gpu_comm_init");
  }
  gpu_check(p, girih_gpu_set_option(p->gpu, "variant", p->gpu_variant), "girih_gpu_set_option");
  gpu_check(p, girih_gpu_set_option(p->gpu, "overlap", p->gpu_overlap), "girih_gpu_set_option");
  {
    /* halo copy / halo push: every rank thread maps its z neighbours' arrays (one process: peer access).  --gpu-copy is
     * on by default (-1) for the halo-first and Diamond steppers; if any GPU cannot map a neighbour the run falls back
     * to the NCCL exchange on every rank -- unless the flag was given explicitly, then it is an error. */
    const int zslabs = p->mpi_size > 1 && p->t.shape[0] == 1 && p->t.shape[1] == 1;
    const int want_copy = p->gpu_copy > 0 || (p->gpu_copy < 0 && !p->gpu_push && p->target_ts >= 1);
    if ((p->gpu_push || want_copy) && zslabs) {
      unsigned char mine[GIRIH_PEER_BLOB_BYTES];
      unsigned char *all = (unsigned char *)malloc((size_t)p->mpi_size * GIRIH_PEER_BLOB_BYTES);
      double ok = 1.0, okmin = 1.0;
      int rc1 = girih_gpu_peer_export(p->gpu, mine, sizeof(mine)), rc2 = GIRIH_OK, rc3 = GIRIH_OK;
      if (rc1 != GIRIH_OK) memset(mine, 0, sizeof(mine));
      team_allgather(mine, sizeof(mine), all, p->mpi_rank);
      if (rc1 == GIRIH_OK && p->mpi_rank > 0)
        rc2 = girih_gpu_peer_attach(p->gpu, 0, all + (size_t)(p->mpi_rank - 1) * GIRIH_PEER_BLOB_BYTES, GIRIH_PEER_BLOB_BYTES);
      if (rc1 == GIRIH_OK && rc2 == GIRIH_OK && p->mpi_rank + 1 < p->mpi_size)
        rc3 = girih_gpu_peer_attach(p->gpu, 1, all + (size_t)(p->mpi_rank + 1) * GIRIH_PEER_BLOB_BYTES, GIRIH_PEER_BLOB_BYTES);
      free(all);
      if (rc1 != GIRIH_OK || rc2 != GIRIH_OK || rc3 != GIRIH_OK) ok = 0.0;
      team_reduce(&ok, NULL, &okmin, NULL, 1, p->mpi_rank);
      team_bcast(&okmin, sizeof(okmin), 0, p->mpi_rank);
      if (okmin > 0.5) {
        gpu_check(p, girih_gpu_set_option(p->gpu, want_copy ? "halo_copy" : "halo_push", 1), "girih_gpu_set_option");
      } else if (p->gpu_push || p->gpu_copy > 0) {
        gpu_check(p, rc1 != GIRIH_OK ? rc1 : (rc2 != GIRIH_OK ? rc2 : (rc3 != GIRIH_OK ? rc3 : GIRIH_ERR_UNSUPPORTED)), "girih_gpu_peer_attach");
      } else {
        girih_gpu_peer_detach(p->gpu);   /* every rank: NCCL exchange */
      }
    }
  }
  gpu_check(p, girih_gpu_set_option(p->gpu, "contract", p->gpu_contract), "girih_gpu_set_option");
  gpu_check(p, girih_gpu_upload(p->gpu, p->U1, p->U2, p->U3, p->coef), "girih_gpu_upload");
  if (p->gpu_tune) {
    /* like the reference's auto-tuner this runs before the measured tests; the fields it advanced are
     * uploaded again.  Every rank tunes its own slab; rank 0 prints. */
    int bt = 0;
    gpu_check(p, girih_gpu_autotune(p->gpu, p->target_ts == 2 && p->gpu_tfuse <= 0, p->mpi_rank == 0 && p->verbose, &bt, NULL, NULL),
              "girih_gpu_autotune");
    if (p->target_ts == 2 && p->gpu_tfuse <= 0) { /* all slabs must run the same pass schedule: rank 0 decides */
      team_bcast(&bt, sizeof(bt), 0, p->mpi_rank);
      p->gpu_tfuse = bt;
    }
    gpu_check(p, girih_gpu_upload(p->gpu, p->U1, p->U2, p->U3, p->coef), "girih_gpu_upload");
  }
}

void gpu_detach(Parameters *p) {
  if (p->gpu) girih_gpu_destroy(p->gpu);
  p->gpu = NULL;
}
