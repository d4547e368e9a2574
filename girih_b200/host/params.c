/*
 * params.c -- defaults, command line, --help / --list, parameter report.
 * Follows src/utils.c:41-123 (param_default), :984-1066 (print_param), :1068-1119 (list_kernels),
 * :1121-1218 (print_help) and :1219-1356 (parse_args) of the reference: same flags, same defaults,
 * same "Key: value" report lines (scripts/parse.py:95-155 keys on them), same exit codes
 * (--help, --list and a bad flag all exit with status 0, src/utils.c:1116-1118,1215-1217,1308-1323).
 */
#define _GNU_SOURCE
#include <getopt.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "girih_host.h"

const char *MWD_name[] = {"Wavefront", "Fixed execution wavefronts", "Relaxed synchronization wavefront",
                          "Relaxed synchronization wavefront with fixed execution", 0};

void reset_timers(Profile *pr) { memset(pr, 0, sizeof(*pr)); }

void param_default(Parameters *p) {
  const int rank = p->mpi_rank, size = p->mpi_size;
  /* src/utils.c:116-117: double literals rounded to real_t */
  static const double coef[11] = {-0.28472, 0.16000, -0.02000, 0.00254, -0.00018, -0.18472,
                                  0.19,     -0.0500, 0.00554,  -0.0009, 0.00354};
  int i;
  memset(p, 0, sizeof(*p));
  p->mpi_rank = rank;
  p->mpi_size = size;
  p->stencil_shape[0] = 256;
  p->stencil_shape[1] = 64;
  p->stencil_shape[2] = 64;
  p->alignment = 8;
  p->target_ts = 0;
  p->target_kernel = 0;
  p->n_tests = 3;
  p->nt = 100;
  p->verbose = 1;
  p->array_padding = 1;
  p->t_dim = -1;
  p->halo_concat = 1;
  p->z_contig = 1;
  p->num_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  p->th_x = p->th_y = p->th_z = p->th_c = -1;
  p->thread_group_size = -1;
  p->wavefront = 1;
  p->num_wf = -1;
  p->t.shape[0] = p->t.shape[1] = p->t.shape[2] = 1;
  p->gpu_overlap = 0;
  p->gpu_push = 0;
  p->gpu_copy = -1;   /* auto: on for --npz > 1 z-slab runs of the halo-first and Diamond steppers when the GPUs can map each other */
  p->gpu_contract = 0;
  p->gpu_tune = 0;
  for (i = 0; i < 11; i++) p->g_coef[i] = (real_t)coef[i];
  reset_timers(&p->prof);
}

static const char *coeff_name(int c) {
  switch (c) {
    case GIRIH_COEF_CONSTANT: return "constant";
    case GIRIH_COEF_VARIABLE: return "variable";
    case GIRIH_COEF_VARIABLE_AXSYM: return "variable axis-symmetric";
    case GIRIH_COEF_VARIABLE_NOSYM: return "variable no-symmetry";
    default: return "Solar kernel";
  }
}

void list_kernels(Parameters *p) {
  int i;
  if (p->mpi_rank == 0) {
    printf("Available time steppers:\n#    Name\n");
    for (i = 0; TSList[i].name != 0; i++) printf("%02d   %s\n", i, TSList[i].name);
    printf("\nAvailable stencil kernels:\n");
    for (i = 0; i < girih_kernel_count(); i++) {
      girih_kernel_desc d;
      girih_kernel_info(i, &d);
      printf("%02d  stencil_op:%s  time-order:%d  radius:%d  coeff:%s\n", i, d.name, d.time_order, d.r,
             coeff_name(d.coeff));
    }
    printf("\nAvailable MWD implementations:\n#    Name\n");
    for (i = 0; MWD_name[i] != 0; i++) printf("%02d   %s\n", i, MWD_name[i]);
  }
  exit(0);
}

void print_help(Parameters *p) {
  if (p->mpi_rank == 0) {
    printf(
        "Note: default values are set in param_default() (girih_b200/host/params.c)\n"
        "Usage:\n\n"
        "  --help\n       Show available options\n"
        "  --list\n       List the time steppers, stencil kernels and MWD variants with their numbers\n"
        "  --verify <bool>\n       Check the selected time stepper against the serial reference kernels\n"
        "       (disables the performance measurement)\n"
        "\nGeneral experiment parameters:\n"
        "  --target-ts <integer>\n       Time stepper (see --list)\n"
        "  --target-kernel <integer>\n       Stencil kernel (see --list)\n"
        "  --nx <integer>  --ny <integer>  --nz <integer>\n       Global domain size\n"
        "  --nt <integer>\n       Number of time steps\n"
        "  --npx <integer>  --npy <integer>  --npz <integer>\n"
        "       Process topology: one GPU per process, npx*npy*npz GPUs.  z-slabs (--npz) are the fast path;\n"
        "       --npx/--npy > 1 is available for the single-step steppers (ts 0, 1), the Diamond stepper needs\n"
        "       --npx 1 --npy 1\n"
        "  --n-tests <integer>\n       Repetitions of the time stepper in a performance run\n"
        "  --alignment <integer>\n       Alignment of the allocated host arrays\n"
        "\nDisplay options:\n"
        "  --verbose <bool>\n       Print the configuration\n"
        "  --debug <bool>\n       Print decomposition details\n"
        "\nSpecialized arguments:\n"
        "  --t-dim <integer>    (Diamond stepper)\n       Time unroll of the diamond; fixes the rounding of --nt\n"
        "  --mwd-type <int>\n       MWD variant (all variants map to the one fused GPU sweep)\n"
        "  --gpu-tfuse <int>\n       Time steps fused per HBM pass by the Diamond stepper (0 = default)\n"
        "  --gpu-variant <int>\n       0 streamed kernels (default), 1 naive kernels\n"
        "  --gpu-tune <bool>\n       Measure every fusion depth and tile shape of the selected operator on the device before the\n       run and use the fastest (printed under [AUTO TUNE]; default 0 = built-in defaults)\n"
        "  --gpu-contract <bool>\n       Evaluate the stencil with the fused multiply-adds gcc emits for the reference under -O3 -mfma\n       (bit-identical to the reference built that way; default 0 = no FMA, bit-identical to the\n       reference's -O0 verifier).  --verify then passes on relative Linf <= 1e-12 (fp64) / 1e-5 (fp32)\n"
        "  --gpu-copy <bool>\n       --npz > 1: outer parts of every slab first, then the halos travel into the neighbouring GPUs' halo planes by\n       copy engine (peer memory) under the sweep of the inner part (default: on for --target-ts 1 and 2 when the GPUs\n       can map each other's memory, else the NCCL exchange)\n"
        "  --gpu-push <bool>\n       Diamond stepper, --npz > 1, 7-point constant operator: the fused sweep stores its boundary planes\n       straight into the neighbouring GPUs' halos over NVLink (no exchange between passes)\n"
        "  --gpu-overlap <bool>\n       Diamond stepper: compute the slab boundaries first and overlap the deep-halo exchange\n       with the interior (default 0: one blocking exchange per fused pass measured faster)\n"
        "  --z-mpi-contig <bool>  --halo-concatenate <integer>  --thread-group-size <integer>\n"
        "  --thx/--thy/--thz/--thc <integer>  --cache-size <integer>  --wavefront <bool>\n"
        "  --num-wavefronts <int>  --use-omp-stat-sched  --threads n[:block[:stride]]\n"
        "  --pad-array  --disable-source-point\n"
        "       CPU tuning flags of the reference: accepted for command-line compatibility,\n"
        "       validated like the reference where they affect results, otherwise ignored\n");
  }
  exit(0);
}

void parse_args(int argc, char **argv, Parameters *p) {
  static struct option long_options[] = {
      {"nz", 1, 0, 0}, {"ny", 1, 0, 0}, {"nx", 1, 0, 0}, {"nt", 1, 0, 0}, {"alignment", 1, 0, 0},
      {"verbose", 1, 0, 0}, {"debug", 1, 0, 0}, {"target-ts", 1, 0, 0}, {"target-kernel", 1, 0, 0},
      {"n-tests", 1, 0, 0}, {"verify", 1, 0, 0}, {"list", 0, 0, 0}, {"help", 0, 0, 0}, {"npx", 1, 0, 0},
      {"npy", 1, 0, 0}, {"npz", 1, 0, 0}, {"t-dim", 1, 0, 0}, {"z-mpi-contig", 1, 0, 0},
      {"disable-source-point", 0, 0, 0}, {"halo-concatenate", 1, 0, 0}, {"thread-group-size", 1, 0, 0},
      {"cache-size", 1, 0, 0}, {"wavefront", 1, 0, 0}, {"num-wavefronts", 1, 0, 0}, {"pad-array", 0, 0, 0},
      {"mwd-type", 1, 0, 0}, {"thx", 1, 0, 0}, {"thy", 1, 0, 0}, {"thz", 1, 0, 0}, {"thc", 1, 0, 0},
      {"threads", 1, 0, 0}, {"use-omp-stat-sched", 0, 0, 0},
      {"gpu-tfuse", 1, 0, 0}, {"gpu-variant", 1, 0, 0}, {"gpu-overlap", 1, 0, 0}, {"gpu-push", 1, 0, 0}, {"gpu-copy", 1, 0, 0}, {"gpu-contract", 1, 0, 0}, {"gpu-tune", 1, 0, 0},
      {0, 0, 0, 0}};
  int c, cache_size = -1;
  optind = 1;
  while (1) {
    int oi = 0;
    const char *n;
    c = getopt_long(argc, argv, "", long_options, &oi);
    if (c == -1) break;
    if (c != 0) {
      if (p->mpi_rank == 0) fprintf(stderr, "Invalid arguments\n\n");
      print_help(p);
    }
    n = long_options[oi].name;
#define IS(s) (strcmp(n, s) == 0)
    if (IS("nz")) p->stencil_shape[2] = atoi(optarg);
    else if (IS("ny")) p->stencil_shape[1] = atoi(optarg);
    else if (IS("nx")) p->stencil_shape[0] = atoi(optarg);
    else if (IS("nt")) p->nt = atoi(optarg);
    else if (IS("npx")) p->t.shape[0] = atoi(optarg);
    else if (IS("npy")) p->t.shape[1] = atoi(optarg);
    else if (IS("npz")) p->t.shape[2] = atoi(optarg);
    else if (IS("alignment")) p->alignment = atoi(optarg);
    else if (IS("verbose")) p->verbose = atoi(optarg) != 0;
    else if (IS("target-ts")) p->target_ts = atoi(optarg);
    else if (IS("target-kernel")) p->target_kernel = atoi(optarg);
    else if (IS("n-tests")) p->n_tests = atoi(optarg);
    else if (IS("verify")) p->verify = atoi(optarg) != 0;
    else if (IS("debug")) p->debug = atoi(optarg) != 0;
    else if (IS("t-dim")) p->t_dim = atoi(optarg);
    else if (IS("z-mpi-contig")) p->z_contig = atoi(optarg) != 0;
    else if (IS("list")) list_kernels(p);
    else if (IS("help")) print_help(p);
    else if (IS("disable-source-point")) { /* source point updates are always off, src/utils.c:336 */ }
    else if (IS("halo-concatenate")) p->halo_concat = atoi(optarg) != 0;
    else if (IS("thread-group-size")) p->thread_group_size = atoi(optarg);
    else if (IS("cache-size")) cache_size = atoi(optarg);
    else if (IS("wavefront")) p->wavefront = atoi(optarg) != 0;
    else if (IS("num-wavefronts")) p->num_wf = atoi(optarg);
    else if (IS("pad-array")) p->array_padding = 1;
    else if (IS("mwd-type")) p->mwd_type = atoi(optarg);
    else if (IS("use-omp-stat-sched")) p->use_omp_stat_sched = 1;
    else if (IS("thx")) p->th_x = atoi(optarg);
    else if (IS("thy")) p->th_y = atoi(optarg);
    else if (IS("thz")) p->th_z = atoi(optarg);
    else if (IS("thc")) p->th_c = atoi(optarg);
    else if (IS("threads")) { /* CPU affinity control: no meaning on the GPU */ }
    else if (IS("gpu-tfuse")) p->gpu_tfuse = atoi(optarg);
    else if (IS("gpu-variant")) p->gpu_variant = atoi(optarg);
    else if (IS("gpu-overlap")) p->gpu_overlap = atoi(optarg) != 0;
    else if (IS("gpu-push")) p->gpu_push = atoi(optarg) != 0;
    else if (IS("gpu-copy")) p->gpu_copy = atoi(optarg) != 0;
    else if (IS("gpu-contract")) p->gpu_contract = atoi(optarg) != 0;
    else if (IS("gpu-tune")) p->gpu_tune = atoi(optarg) != 0;
#undef IS
  }
  if (optind < argc) {
    if (p->mpi_rank == 0) fprintf(stderr, "Invalid arguments\n\n");
    print_help(p);
  }
  if (cache_size != -1) p->cache_size = cache_size;
  p->orig_thread_group_size = p->thread_group_size;
}

void print_param(const Parameters *p) {
  const char *precision = (sizeof(real_t) == 4) ? "SP" : "DP";
  printf("\n******************************************************\n");
  printf("Parameters settings\n");
  printf("******************************************************\n");
  printf("Time stepper name: %s\n", TSList[p->target_ts].name);
  printf("Stencil Kernel name: %s\n", p->stencil.name);
  printf("Stencil Kernel semi-bandwidth: %d\n", p->stencil.r);
  printf("Stencil Kernel coefficients: %s\n", coeff_name(p->stencil.coeff));
  printf("Precision: %s\n", precision);
  printf("Global domain    size:%llu    nx:%d    ny:%d    nz:%d\n", (unsigned long long)p->n_stencils,
         p->stencil_shape[0], p->stencil_shape[1], p->stencil_shape[2]);
  printf("Rank 0 domain    size:%llu    nx:%d    ny:%d    nz:%d\n", (unsigned long long)p->ln_stencils,
         p->lstencil_shape[0], p->lstencil_shape[1], p->lstencil_shape[2]);
  printf("Number of time steps: %d\n", p->nt);
  printf("Alignment size: %d Bytes\n", p->alignment);
  printf("Number of tests: %d\n", p->n_tests);
  printf("Verify:   %d\n", p->verify);
  printf("Source point enabled: %d\n", 0);
  printf("Time unroll:   %d\n", p->t_dim);
  printf("Using separate call to central line update: %d\n", 0);
  printf("Halo concatenation: %d\n", p->halo_concat);
  switch (p->target_ts) {
    case 0:
    case 1:
      if (p->z_contig == 1) printf("MPI datatype is contiguous across the Z direction\n");
      printf("Block size in Y: %d\n", p->ldomain_shape[1]);
      printf("OpenMP schedule: %s\n", p->use_omp_stat_sched ? "static" : "static1");
      break;
    case 2:
      printf("Enable wavefronts: %d\n", p->wavefront != 0);
      printf("Wavefront parallel strategy: %s\n", MWD_name[p->mwd_type]);
      printf("Intra-diamond width:   %d\n", (p->t_dim + 1) * 2 * p->stencil.r);
      printf("Wavefront width:  %d\n", (p->t_dim * 2) * p->stencil.r + p->num_wf);
      printf("Intra-diamond prologue/epilogue MStencils: %llu\n",
             (unsigned long long)(p->idiamond_pro_epi_logue_updates / (1000 * 1000)));
      printf("Multi-wavefront updates: %d\n", p->num_wf);
      printf("User set thread group size: %d\n", p->orig_thread_group_size);
      printf("Thread group size: %d\n", p->thread_group_size);
      printf("GPU fused time steps per pass: %d\n", p->gpu_tfuse);
      break;
  }
  printf("OpenMP Threads: %d\n", p->num_threads);
  printf("Assumed usable cache size: %dKiB\n", p->cache_size);
  printf("MPI size: %d\n", p->mpi_size);
  printf("Processors topology (npx, npy, npz): %02d,%02d,%02d\n", p->t.shape[0], p->t.shape[1], p->t.shape[2]);
  printf("GPU kernels: %s\n", p->gpu_variant == 1 ? "naive" : "streamed (sm_100a)");
  printf("GPU arithmetic: %s\n", p->gpu_contract ? "FMA-contracted (reference built with -mfma)" : "separate multiply/add (reference verifier)");
  printf("******************************************************\n");
}
