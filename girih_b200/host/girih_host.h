/*
 * girih_host.h -- host side of the B200 build of GIRIH's `mwd_kernel`: parameters, the operator and
 * time-stepper tables, initialisation, the timing harness and the verifier.  Plain C (gnu99), no
 * CUDA types: everything device-side is reached through include/girih_cuda.h.
 *
 * It mirrors the reference's host interface (names, argument meaning, error behaviour):
 *   Parameters / struct Stencil / struct time_stepper   src/data_structures.h:202-297
 *   param_default, parse_args, init, arrays_allocate, init_coeff, domain_data_fill,
 *   print_param, list_kernels, performance_results      src/utils.c
 *   performance_test                                    src/performance.c:29-127
 *   verify, verify_serial_generic, compare_results_std  src/verification.c
 *   TSList[]                                            src/wrappers.h:29-37
 * Precision is a compile-time switch like the reference (-DDP=1, src/data_structures.h:90-105):
 * build/mwd_kernel is fp32, build_dp/mwd_kernel is fp64.
 *
 * MPI ranks of the reference become "ranks" that are host threads of ONE process, one per GPU
 * (--npz N); the team_* helpers stand in for MPI_Barrier / MPI_Reduce / the verification gather.
 */
#ifndef GIRIH_HOST_H_
#define GIRIH_HOST_H_

#include <stdint.h>
#include <stdio.h>

#include "../../include/girih_cuda.h"

#ifndef DP
#define DP 0
#endif
#if DP
typedef double real_t;
#else
typedef float real_t;
#endif

#define BOUNDARY_SRC_VAL (100.1)   /* src/data_structures.h:77 */

typedef struct {
  double compute, communicate, send_recv, wait, total, others, ts_main, ts_others;
} Profile;                          /* src/data_structures.h:141-143 */

struct Stencil {                    /* src/data_structures.h:211-223, GPU-relevant part */
  const char *name;
  int r, time_order, nd, shape, coeff;
};

typedef struct {
  int shape[3];                     /* --npx --npy --npz */
  int rank_coords[3];
} Topology;

typedef struct Parameters {
  /* experiment */
  int alignment, verbose, debug, verify, n_tests, nt;
  int stencil_shape[3];             /* global interior nx, ny, nz */
  int target_ts, target_kernel, mwd_type;
  int t_dim, wavefront, num_wf, thread_group_size, th_x, th_y, th_z, th_c;
  int cache_size, halo_concat, array_padding, use_omp_stat_sched, z_contig;
  int num_threads;                  /* host threads available (reported only) */
  int orig_thread_group_size;
  /* GPU knobs (new flags, defaults keep the reference CLI valid) */
  int gpu_tfuse;                    /* --gpu-tfuse: fused steps per HBM pass for ts 2 (0 = auto) */
  int gpu_variant;                  /* --gpu-variant: 0 auto, 1 naive kernels */
  int gpu_overlap;                  /* --gpu-overlap: overlap halo exchange with compute */
  int gpu_push;                     /* --gpu-push: fused passes store boundary planes into the neighbours' halos (peer memory) */
  int gpu_copy;                     /* --gpu-copy: overlapped passes, halos moved into the neighbours' halo planes by the copy engines */
  int gpu_tune;                     /* --gpu-tune: on-device search over fusion depth and tiles ([AUTO TUNE]) */
  int gpu_contract;                 /* --gpu-contract: FMA-contracted arithmetic (reference built with -mfma) */
  /* decomposition */
  int mpi_rank, mpi_size;
  Topology t;
  int lstencil_shape[3], ldomain_shape[3], gb[3], ge[3];
  uint64_t n_stencils, ln_domain, ln_stencils;
  uint64_t idiamond_pro_epi_logue_updates;
  /* data */
  struct Stencil stencil;
  real_t g_coef[11];
  real_t *U1, *U2, *U3, *coef;
  /* device */
  girih_gpu_ctx *gpu;
  int gpu_device;
  int steps_executed;               /* steps the last stepper call really ran */
  int tfuse_used;
  Profile prof;
} Parameters;

struct time_stepper {               /* src/data_structures.h:294-297 */
  const char *name;
  void (*func)(Parameters *p);
};
extern struct time_stepper TSList[];
extern const char *MWD_name[];

/* params.c */
void param_default(Parameters *p);
void parse_args(int argc, char **argv, Parameters *p);
void print_help(Parameters *p);
void list_kernels(Parameters *p);
void print_param(const Parameters *p);
void reset_timers(Profile *pr);

/* init.c */
void init(Parameters *p);
void arrays_allocate(Parameters *p);
void arrays_free(Parameters *p);
void init_coeff(Parameters *p);
void domain_data_fill(Parameters *p);
uint64_t coef_array_size(const Parameters *p);
void gpu_attach(Parameters *p);     /* create context, join communicator, upload */
void gpu_detach(Parameters *p);
void girih_fatal(const Parameters *p, const char *fmt, ...);   /* "ERROR: ..." + exit(1) */

/* steppers.c */
void gpu_naive_ts(Parameters *p);
void gpu_halo_first_ts(Parameters *p);
void gpu_diamond_ts(Parameters *p);

/* perf.c */
void performance_test(Parameters *p);

/* verify.c */
void verify(Parameters *p);
int verify_compute(Parameters *p, double *max_err, double *l1_err, double *max_ref);
/* solar.c: table slot 6 (12 complex fields in one array, 28 complex coefficient arrays) */
int is_solar(const Parameters *p);
void solar_init_coeff(Parameters *p);
void solar_domain_fill(Parameters *p);
int solar_verify_compute(Parameters *p, double *max_err, double *l1_err, double *max_ref);

/* team.c -- the ranks of one process */
void team_init(int nranks);
void team_barrier(void);
void team_reduce(const double *in, double *max, double *min, double *sum, int n, int rank);
void team_bcast(void *buf, size_t len, int root, int rank);
void team_allgather(const void *mine, size_t len, void *all, int rank);   /* all: nranks * len bytes, len <= 256 */
void *team_shared_alloc(size_t bytes, int rank);     /* collective: same pointer on every rank */
void team_shared_free(void *ptr, int rank);
void team_run(int nranks, void (*fn)(int rank, void *arg), void *arg);

#endif /* GIRIH_HOST_H_ */
