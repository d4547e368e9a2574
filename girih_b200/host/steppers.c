/*
 * steppers.c -- the time-stepper table.  Same names and order as the reference's TSList[]
 * (src/wrappers.h:29-37), so --target-ts and --list mean the same thing; each entry drives the
 * sm_100a kernels through the C ABI instead of the OpenMP loop nests.
 *
 *   0 "Spatial Blocking"  naive_nonblocking_ts      src/kernels/nb_naive_ts.c:165-216
 *   1 "Halo-first"        halo_first_ts             src/kernels/halo_first_ts.c:196-330
 *   2 "Diamond"           dynamic_intra_diamond_ts  src/kernels/diamond_ts.c:871-976
 *
 * Like the reference's steppers these functions return nothing and terminate the process on error.
 * They leave the result on the device: verify() and performance_test() fetch p->U1 when they need it
 * (girih_gpu_download / girih_gpu_scan_u1), which keeps host<->device copies out of the timed region
 * exactly as the reference keeps allocation and fill out of it (src/performance.c:50-52,70-78).
 */
#include <stdlib.h>

#include "girih_host.h"

struct time_stepper TSList[] = {
    {"Spatial Blocking", gpu_naive_ts},
    {"Halo-first", gpu_halo_first_ts},
    {"Diamond", gpu_diamond_ts},
    {0, 0},
};

static void finish(Parameters *p, int rc, const char *what) {
  double comp = 0, comm = 0, total = 0;
  int nk = 0, np = 0, ns = 0, tf = 1;
  if (rc != GIRIH_OK) {
    fprintf(stderr, "ERROR: %s: %s (%s)\n", what, girih_gpu_strerror(rc), girih_gpu_last_error(p->gpu));
    exit(1);
  }
  girih_gpu_last_elapsed_ms(p->gpu, &comp, &comm, &total);
  girih_gpu_last_launch_info(p->gpu, &nk, &np, &ns, &tf);
  p->steps_executed = ns;
  p->tfuse_used = tf;
  /* Profile fields as the reference fills them (src/kernels/nb_naive_ts.c:201-202): compute = device time inside the
   * sweeps (cudaEvent pairs around every launch), communicate = device time of the exchanges on the comm stream (with
   * the overlapped schedules it runs under the sweeps, so the two may add up to more than the total), wait = what is
   * left of the total: the compute stream waiting for an exchange or for a neighbour */
  p->prof.compute += 1e-3 * comp;
  p->prof.communicate += 1e-3 * comm;
  p->prof.wait += 1e-3 * (total - comp > 0 ? total - comp : 0);
  p->prof.total = 1e-3 * total;
}

/* U1 <- step(U2); U2 <- step(U1); nt/2 times: an odd nt executes nt+1 steps (nb_naive_ts.c:187) */
static int naive_steps(const Parameters *p) { return (p->nt + 1) / 2 * 2; }

void gpu_naive_ts(Parameters *p) { finish(p, girih_gpu_run_single(p->gpu, naive_steps(p), 0), "girih_gpu_run_single"); }

void gpu_halo_first_ts(Parameters *p) { finish(p, girih_gpu_run_single(p->gpu, naive_steps(p), 1), "girih_gpu_run_single"); }

/* The diamond stepper executes nt-1 steps after its rounding of nt (derived from
 * diamond_ts.c:874-877,767,642,836), leaving U1 = level nt-1 and U2 = level nt-2. */
void gpu_diamond_ts(Parameters *p) {
  finish(p, girih_gpu_run_fused(p->gpu, p->nt - 1, p->gpu_tfuse), "girih_gpu_run_fused");
  p->prof.ts_main += p->prof.total;
}
