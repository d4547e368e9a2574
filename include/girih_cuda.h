/*
 * girih_cuda.h -- C ABI of the B200 (sm_100a) implementation of GIRIH's star-stencil time
 * stepper.  Plain C: pointers, ints and sizes only; nothing here throws, exits or prints.
 *
 * What it replaces in the reference (paths relative to the GIRIH source tree):
 *   - the operator plugin point  spt_blk_func_t / mwd_func_t   src/data_structures.h:202-209,
 *     tables src/kernels/stencils.c:260-352, selected by set_kernels() src/utils.c:253-300
 *   - the time-stepper plugin point  struct time_stepper {name, void (*func)(Parameters*)}
 *     src/data_structures.h:294-297, table TSList[] src/wrappers.h:29-37, called from
 *     src/performance.c:75 and src/verification.c:37
 *   - the halo exchange of src/mpi_utils.c:116-202 + src/kernels/nb_naive_ts.c:32-155
 *     (z-slab decomposition only, NCCL send/recv instead of MPI)
 * The C host (girih_b200/host/, the `mwd_kernel` executable) registers three steppers in its own
 * TSList[] that call girih_gpu_run_single / girih_gpu_run_fused below; INTEGRATION.md shows the
 * stub a GIRIH maintainer would add to the original tree.
 *
 * Ownership: the caller owns the host arrays (Parameters.U1/U2/U3/coef, allocated by
 * arrays_allocate(), src/utils.c:153-218) before and after every call; the context owns all
 * device memory.  After a stepper returns, the reference reads p->U1 (src/verification.c:974,
 * src/utils.c:820-828): call girih_gpu_download() for that.
 *
 * Error convention: every function returns 0 (GIRIH_OK) or a positive girih_status; the C host
 * turns non-zero into the reference's "ERROR: ..." + exit(1) (src/data_structures.h:299-314).
 * There is NO CPU fallback: without a CUDA device girih_gpu_create() fails with
 * GIRIH_ERR_NO_DEVICE.
 *
 * Threading: a context is bound to one CUDA device and must be driven by one host thread at a
 * time.  Multi-GPU = one context per GPU (one process per GPU under torchrun, or one host thread
 * per GPU inside mwd_kernel --npz N), joined by girih_gpu_comm_init().
 */
#ifndef GIRIH_CUDA_H_
#define GIRIH_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct girih_gpu_ctx girih_gpu_ctx;

enum girih_status {
  GIRIH_OK = 0,
  GIRIH_ERR_ARG = 1,          /* bad argument / shape                                   */
  GIRIH_ERR_NO_DEVICE = 2,    /* no CUDA device or driver: there is no CPU fallback     */
  GIRIH_ERR_CUDA = 3,         /* a CUDA runtime call failed (see girih_gpu_last_error)  */
  GIRIH_ERR_UNSUPPORTED = 4,  /* configuration the selected operator does not offer      */
  GIRIH_ERR_NCCL = 5,         /* NCCL missing or a NCCL call failed                     */
  GIRIH_ERR_STATE = 6,        /* call order (e.g. run before upload)                    */
  GIRIH_ERR_FRAME = 7         /* fused stepping needs identical Dirichlet frames in U1/U2 */
};

/* Stencil_Shapes / Stencil_Coefficients, src/data_structures.h:114-126 (same numbering). */
enum girih_shape { GIRIH_STAR = 0, GIRIH_TTI = 1, GIRIH_BOX = 2 };
enum girih_coeff {
  GIRIH_COEF_CONSTANT = 0, GIRIH_COEF_VARIABLE = 1, GIRIH_COEF_VARIABLE_AXSYM = 2,
  GIRIH_COEF_VARIABLE_NOSYM = 3, GIRIH_COEF_SOLAR = 4
};

/* One row of the operator table: struct StencilInfo, src/data_structures.h:225-233, plus the
 * bookkeeping the GPU side adds. */
typedef struct {
  const char *name;      /* "star" / "box"                                              */
  int r;                 /* semi-bandwidth                                              */
  int time_order;        /* 1 or 2                                                      */
  int nd;                /* domain-sized arrays streamed per step (reference's count)   */
  int shape;             /* enum girih_shape                                            */
  int coeff;             /* enum girih_coeff                                            */
  int n_coef_arrays;     /* domain-sized coefficient arrays (0 for constant)            */
  int n_coef_scalars;    /* scalar coefficients read from coef[] (constant kernels)     */
  int words_per_lup;     /* algorithmic words moved per lattice update      */
  int max_tfuse;         /* deepest temporal fusion the GPU stepper offers (>=1)        */
  int gpu_supported;     /* 1 for every slot since round 2 (0 = src/kernels/stencils.h:40-47 analogue) */
} girih_kernel_desc;

/* stencil_info_list[] of src/kernels/stencils.c:260-271: indices 0..7 as printed by --list.
 *
 * Slot 6 ("solar", src/kernels/solar_spt_blk.ic:20-397) differs from the star / box operators in its data, exactly as in
 * the reference: ONE field array of 12 complex components, u[12][nnz][nny][nnx][re,im] (src/utils.c:168-172; U2 == 0),
 * and 28 complex coefficient arrays of the same shape (src/utils.c:199-201), no x padding (src/utils.c:359-361).
 *   girih_gpu_create      domain_shape = stencil_shape + 2 in every direction; nranks must be 1 (the reference's halo
 *                         exchange moves one real per cell, src/mpi_utils.c:173-200)
 *   girih_gpu_upload      U1 = the field array (24 reals per cell), coef = the 56 reals per cell; U2 / U3 ignored (NULL)
 *   girih_gpu_download    U1 only
 *   girih_gpu_run_single  nsteps time steps; a time step is the H update of every interior cell followed by the E update
 *                         (solar(), solar_spt_blk.ic:388-397, ALL_FIELDS), two kernel launches, in place
 *   girih_gpu_step_box    solar() over a box of cells (dst ignored)
 *   girih_gpu_run_fused, option contract, pipelined transfers: GIRIH_ERR_UNSUPPORTED (the reference's default wavefront
 *                         table holds not_supported_mwd for this slot, src/kernels/stencils.c:303-311)
 * words_per_lup = 104: 24 field reals read, 24 written, 56 coefficient reals read per cell and time step. */
int girih_kernel_count(void);
int girih_kernel_info(int target_kernel, girih_kernel_desc *out);

/* Number of visible CUDA devices (0 and GIRIH_ERR_NO_DEVICE when there is none). */
int girih_gpu_count(int *n);

/*
 * Create the device-side state of ONE z-slab.
 *   device          CUDA ordinal
 *   target_kernel   operator table index (--target-kernel)
 *   elem_size       4 (float build) or 8 (-DDP=1 build), src/data_structures.h:90-105
 *   stencil_shape   LOCAL interior nx,ny,nz of this slab (Parameters.lstencil_shape)
 *   domain_shape    LOCAL host array shape nnx,nny,nnz = interior + 2r (+ x padding), x fastest
 *                   (Parameters.ldomain_shape, src/utils.c:367-374); z halo planes between slabs
 *                   are r deep on the host side exactly as in src/mpi_utils.c:173-200
 *   rank, nranks    position in the 1-D z chain (--npz); rank 0 holds the lowest z
 */
int girih_gpu_create(girih_gpu_ctx **ctx, int device, int target_kernel, int elem_size,
                     const int stencil_shape[3], const int domain_shape[3], int rank, int nranks);
void girih_gpu_destroy(girih_gpu_ctx *ctx);

/* Communicator bootstrap for nranks > 1 (replaces mpi_topology_init, src/mpi_utils.c:63-113):
 * rank 0 calls girih_gpu_comm_unique_id(), the host distributes the 128 bytes by its own means
 * (torch.distributed broadcast, a shared variable between threads, ...), every rank calls
 * girih_gpu_comm_init().  NCCL is loaded at run time (libnccl.so.2). */
#define GIRIH_COMM_ID_BYTES 128
int girih_gpu_comm_unique_id(void *id, size_t len);
int girih_gpu_comm_init(girih_gpu_ctx *ctx, const void *id, size_t len);

/* Halo push: compute and z exchange in one kernel over NVLink peer memory (no reference counterpart; the GPU form of
 * the reference's compute/communication overlap, src/kernels/halo_first_ts.c:156-194).  Every rank exports handles of
 * its field arrays and of a flag word (girih_gpu_peer_export, GIRIH_PEER_BLOB_BYTES), the host hands each blob to the
 * rank's two z neighbours, which map it (girih_gpu_peer_attach: which = 0 for the blob of the lower neighbour, 1 for
 * the upper one; rank threads of one process use peer access, separate processes CUDA IPC).  With
 * girih_gpu_set_option("halo_push", 1) the fused passes of slot 1 then store their boundary planes straight into the
 * neighbours' halo planes while they sweep; a pass starts when both neighbours have flagged the end of the previous
 * one (device-side flags, no host synchronisation, no NCCL kernel between passes).  NCCL still carries the first and
 * the last exchange of a run.  Every rank must make the same calls.
 * Halo copy (round 2, the schedule bench.py --gpus N measures): girih_gpu_set_option("halo_copy", 1) with the same
 * mappings keeps the order of halo_first_ts.c:156-194 -- the outer parts of the slab first, exchange, inner part --
 * and lets the COPY ENGINES move the halos: cudaMemcpyAsync from this rank's top / bottom planes into the
 * neighbours' halo planes on the comm stream plus a flag word, under the sweep of the inner part (an NCCL kernel
 * in that place takes SMs from a sweep that fills whole waves).  Every operator and both steppers; z-slabs only. */
#define GIRIH_PEER_BLOB_BYTES 256
int girih_gpu_peer_export(girih_gpu_ctx *ctx, void *blob, size_t len);
int girih_gpu_peer_attach(girih_gpu_ctx *ctx, int which, const void *blob, size_t len);
/* unmaps both neighbours; between processes: detach on every rank, synchronise the ranks, then destroy */
int girih_gpu_peer_detach(girih_gpu_ctx *ctx);

/* Process topology (--npx/--npy/--npz; MPI_Cart_create / MPI_Cart_coords / MPI_Cart_shift of
 * src/mpi_utils.c:63-81): dims = (npx, npy, npz) with npx*npy*npz == nranks, coords = this rank's position,
 * rank == (coords[0]*npy + coords[1])*npz + coords[2] (MPI's row-major order).  Optional -- without it the ranks
 * form z-slabs -- and to be called before girih_gpu_comm_init.  With npx or npy > 1 the steppers also exchange the
 * r-deep x and y faces of src/mpi_utils.c:116-170 after every step (device pack / ncclSend+ncclRecv / unpack, the
 * analogue of sub_array_copy, src/mpi_utils.c:31-45) and run single steps only: like the reference, whose
 * diamond stepper rejects an x decomposition (src/kernels/diamond_utils.c:1035-1040), temporal fusion needs
 * npx == npy == 1 here. */
int girih_gpu_set_topology(girih_gpu_ctx *ctx, const int dims[3], const int coords[3]);

/* Host -> device of the arrays arrays_allocate()/init_coeff()/domain_data_fill() produced
 * (src/performance.c:50-52).  U3 (roc2) may be NULL unless time_order == 2; coef holds
 * n_coef_scalars values (constant) or n_coef_arrays*ln_domain values (variable), in the
 * reference's layout COEF(m,i,j,k) = coef[idx + ln_domain*m], src/kernels/stencils.h:32.
 * Outside the timed region, like the reference's allocation/fill. */
int girih_gpu_upload(girih_gpu_ctx *ctx, const void *U1, const void *U2, const void *U3,
                     const void *coef);
/* Device -> host of U1 and/or U2 (either may be NULL) into arrays of domain_shape. */
int girih_gpu_download(girih_gpu_ctx *ctx, void *U1, void *U2);
/* Same transfers from/to page-locked host memory, asynchronous on the context's stream and
 * followed by a stream synchronise -- used for end-to-end timing. */
int girih_gpu_upload_fields(girih_gpu_ctx *ctx, const void *U1, const void *U2);
/* Pipelined transfers for a stream of independent jobs (performance_test()'s n_tests loop with fresh inputs, a
 * parameter sweep, shots of a survey): the host -> device copy of job i+1 and the device -> host copy of job i-1
 * run on their own streams (the two copy engines) underneath the sweeps of job i.
 *   prefetch_fields  asynchronous copy of U1 and/or U2 (page-locked, either may be NULL) into device staging; the
 *                    host arrays must stay valid until the next stepper call or girih_gpu_sync_transfers returns
 *   commit_fields    the prefetched fields become the device arrays (stream-ordered behind the copy)
 *   download_async   like girih_gpu_download, but returns once the copy is enqueued
 *   sync_transfers   blocks until every transfer issued so far has completed
 * Typical loop: prefetch(job 0); for i: commit(); prefetch(job i+1); run; download_async(out i); sync_transfers(). */
int girih_gpu_prefetch_fields(girih_gpu_ctx *ctx, const void *U1, const void *U2);
int girih_gpu_commit_fields(girih_gpu_ctx *ctx);
int girih_gpu_download_async(girih_gpu_ctx *ctx, void *U1, void *U2);
int girih_gpu_sync_transfers(girih_gpu_ctx *ctx);

/*
 * Time steppers.  All follow the reference's parity convention: global step s = 1,2,... reads
 * the array written by step s-1 and writes U1 when s is odd, U2 when s is even
 * (src/kernels/nb_naive_ts.c:187-203, src/kernels/diamond_ts.c:440-444).  After `nsteps` steps
 * the newest level is in U1 (nsteps odd) or U2 (even) and the other array holds level nsteps-1,
 * bit-identical to the reference steppers.
 *
 * girih_gpu_run_single: ts 0 "Spatial Blocking" (overlap = 0) and ts 1 "Halo-first"
 *   (overlap = 1: boundary slabs first, halo exchange overlapped with the interior;
 *   src/kernels/halo_first_ts.c:156-194).  One HBM pass per step.
 * girih_gpu_run_fused: ts 2 "Diamond": the GPU analogue of the MWD sweep
 *   (src/kernels/diamond_ts.c:871-976, stencils_1wf.ic:33-82): tfuse time steps are fused per
 *   HBM pass with a z-streaming wavefront; halo depth between slabs = tfuse * r.
 *   tfuse <= 0 selects the default for the operator; tfuse is clamped to max_tfuse.
 * The GIRIH CLI passes nsteps = nt rounded up to even (ts 0/1) or nt-1 after the diamond
 * rounding of nt (ts 2, src/kernels/diamond_utils.c:1042-1056).
 */
int girih_gpu_run_single(girih_gpu_ctx *ctx, int nsteps, int overlap);
int girih_gpu_run_fused(girih_gpu_ctx *ctx, int nsteps, int tfuse);

/* Host-only planning helpers (no device needed).  They return exactly what girih_gpu_run_fused and the
 * halo exchange execute, so the multi-rank logic can be exercised without GPUs:
 *   girih_plan_fused_passes   steps per pass for nsteps steps at depth tfuse (the last entry is the
 *                             single step that leaves U1/U2 as the reference does); sizes may be NULL
 *   girih_plan_halo_exchange  first plane of the four `depth`-plane blocks of one z exchange, in local
 *                             plane coordinates (0 = first interior plane); -1 / <-depth = no neighbour.
 *                             Geometry of src/mpi_utils.c:173-200 with depth = T*r instead of r.
 *   girih_plan_fused_exchanges  halo planes exchanged before each pass of that schedule (0 = none) when one
 *                             exchange serves up to `group` passes and at most halo_cap planes: the first
 *                             pass of a group also sweeps the neighbour's planes the later ones read. */
int girih_plan_fused_passes(int nsteps, int tfuse, int *sizes, int max_sizes, int *n_sizes);
int girih_plan_fused_exchanges(int nsteps, int tfuse, int r, int halo_cap, int group, int *depth, int max_n, int *n);
int girih_plan_halo_exchange(int nz, int depth, int rank, int nranks, int *send_down, int *recv_down,
                             int *send_up, int *recv_up);

/* One application of the operator over the box [xb,xe) x [yb,ye) x [zb,ze) in HOST index space
 * (the spt_blk_func_t contract, src/kernels/stencils_spt_blk.ic:19-48): dst=1 writes U1 from U2,
 * dst=2 writes U2 from U1. */
int girih_gpu_step_box(girih_gpu_ctx *ctx, int dst, int xb, int yb, int zb, int xe, int ye, int ze);

/* cudaEvent timing of the last run_* call, in milliseconds (replaces the MPI_Wtime brackets of
 * src/performance.c:70-78 and the Profile fields of src/data_structures.h:141-143).
 * compute = kernel time on the compute stream, comm = halo exchange on the comm stream,
 * total = first launch to last completion. */
int girih_gpu_last_elapsed_ms(girih_gpu_ctx *ctx, double *compute_ms, double *comm_ms,
                              double *total_ms);
/* Launch accounting of the last run_* call: kernels launched, fused passes, steps executed. */
int girih_gpu_last_launch_info(girih_gpu_ctx *ctx, int *n_kernels, int *n_passes, int *n_steps,
                               int *tfuse_used);
/* Named counters of the context (since creation): "exact_launches" = fused passes that ran on exact (non-overlapping)
 * tiles with edge hand-off between co-resident CTAs, the GPU form of the reference's non-redundant diamond tiles
 * (src/kernels/diamond_ts.c:565-595); "fused_launches" = all fused passes.  GIRIH_ERR_ARG for an unknown key. */
int girih_gpu_get_stat(girih_gpu_ctx *ctx, const char *key, long long *value);

/* Measurement hook for the roofline figure: launches `reps` passes of `tfuse` fused steps back to
 * back over the whole slab (ping-ponging U1/U2, so the fields keep evolving) and returns the
 * average device time of ONE pass = one kernel launch, from cudaEvents on the launching stream.
 * No halo exchange: single-slab contexts only. */
int girih_gpu_time_pass(girih_gpu_ctx *ctx, int tfuse, int reps, double *ms_per_pass);

/* NaN/Inf and near-zero scan of the final U1 over the whole local domain
 * (src/utils.c:819-840), done on the device. */
int girih_gpu_scan_u1(girih_gpu_ctx *ctx, uint64_t *n_nan_inf, uint64_t *n_zero);

/* Tuning knobs (all optional).  Keys:
 *   "variant"  0 auto (marching kernel for single steps, fused sweep for T > 1), 1 naive kernels,
 *              2 fused-sweep kernel also for T = 1
 *   "zchunk"   output planes per CTA (0 = choose)
 *   "tile"     fused sweep: PY*100 + NW (rows per thread, warps per CTA); marching kernel: rows per CTA.
 *              Schedule variants of the tiles 408 and 216, all bit-identical in their results: 9000 + tile
 *              trapezoid skip (warps without a core row skip the last fused level; slots 1, 2, 3, 5), and for
 *              slot 1 5000 + tile split (arrive/wait) CTA barrier, 7000 + tile decoupled levels
 *   "overlap"  fused passes: compute the slab boundaries first and overlap the deep-halo exchange with
 *              the interior (default 0: one exchange per pass, ordered before it, measured faster)
 *   "halo_group" z-slab runs of first-order-in-time operators: fused passes served by one halo exchange
 *              (0 = choose: up to 4 while the recomputed planes stay below 1/16 of the thinnest slab)
 *   "halo_copy" / "halo_push"  see girih_gpu_peer_export above
 *   "zwave", "zwave_block"  operators without a fused-sweep kernel (radius 4, box), one slab: run_fused keeps
 *              `zwave` time steps in flight along z, each r planes behind the previous one, in blocks of
 *              `zwave_block` planes (the reference's wavefront, src/kernels/stencils_1wf.ic:37-77, with the L2 as
 *              the cache).  Default 0 = plain single steps: measured no faster on B200.
 *   "tile" 10408 / 10216 (slot 1, fp64, depth >= 3): exact, non-overlapping tiles whose rims travel between
 *              co-resident CTAs through L2 (kernels_r1x.cuh; falls back to the overlapped tiles when the plane
 *              needs more tiles than the device has SMs)
 *   "contract" arithmetic of the per-point expression.  0 (default): every product and sum rounded
 *              separately -- bit-identical to the reference built without FMA (conf/make.conf.gcc, `-O3`)
 *              and to its -O0 verifier (src/verification.c).  1: the fused multiply-adds gcc emits for the
 *              same FUNC_BODY under `-O3 -mfma` (first product fused onto the second, later products
 *              onto the running sum) -- bit-identical to the reference built that way; 30 % fewer FP64
 *              instructions per lattice update.  With 1 only the default tile and slot 1's schedule variants
 *              (5xxx, 7xxx, 9xxx) exist; any other "tile" runs the default tile. */
int girih_gpu_set_option(girih_gpu_ctx *ctx, const char *key, int value);

/* On-device tuner -- the GPU analogue of auto_tune_params() (src/kernels/diamond_utils.c:691-847, the
 * "measure every feasible blocking, keep the fastest" loop of run_tuning_test, :244-269): times every
 * (fused steps per pass, tile) this operator has kernels for on the resident slab and keeps the fastest for
 * later run_single / run_fused(tfuse = 0) calls of this context.  fused = 0 tunes the single-step pass only.
 * verbose = 1 prints the candidates under the reference's "[AUTO TUNE]" prefix on stdout.  The fields keep
 * evolving while it measures: upload them again afterwards.  Outputs may be NULL. */
int girih_gpu_autotune(girih_gpu_ctx *ctx, int fused, int verbose, int *best_tfuse, int *best_tile,
                       double *best_mlups);

const char *girih_gpu_strerror(int status);
/* Detail of the last failure on this context (CUDA/NCCL error string); never NULL. */
const char *girih_gpu_last_error(girih_gpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* GIRIH_CUDA_H_ */
