/*
 * ref_dump -- runs the UNMODIFIED GIRIH reference time steppers (linked from the
 * objects compiled out of /root/reference/src, see oracle/Makefile) and writes
 * the resulting p.U1 array to a raw file, so the oracle restatement and the CUDA
 * path can be pinned against real reference output.
 *
 * TEST INFRASTRUCTURE -- not product code.
 *
 * This file replaces only the reference's src/driver.c:21-82 `main`: the call
 * sequence below is the same one (param_default -> parse_args ->
 * mpi_topology_init -> init) followed by the body of performance_test()
 * (src/performance.c:45-52: mpi_halo_init, arrays_allocate, init_coeff,
 * domain_data_fill) and ONE call of TSList[ts].func(p) -- exactly what
 * verify() does at src/verification.c:28-37 -- instead of the timing loop.
 *
 * Usage: GIRIH_REF_DUMP=<file> ref_dump_{sp,dp} <mwd_kernel flags>
 * File layout: 8 x int32 header {magic 0x47495249, sizeof(real_t), nnx, nny, nnz,
 *              r, nt (after the diamond stepper's rounding), target_kernel}
 *              followed by nnx*nny*nnz real_t values of U1 (x fastest); for the solar slot
 *              (table index 6, one array of 12 complex fields, src/utils.c:168-172) 24 times as many.
 */
#include "driver.h"
#include <stdint.h>
#include <stdlib.h>

extern void mpi_halo_init(Parameters *);
extern void arrays_allocate(Parameters *);
extern void init_coeff(Parameters *);
extern void domain_data_fill(Parameters *);

int main(int argc, char **argv)
{
  int provided;
  Parameters p;
  const char *path = getenv("GIRIH_REF_DUMP");
  FILE *fp;
  int32_t hdr[8];

  if (path == NULL) { fprintf(stderr, "ref_dump: set GIRIH_REF_DUMP\n"); return 2; }

  MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &provided);
  MPI_Comm_rank(MPI_COMM_WORLD, &(p.mpi_rank));
  MPI_Comm_size(MPI_COMM_WORLD, &(p.mpi_size));
  param_default(&p);
  parse_args(argc, argv, &p);
  mpi_topology_init(&p);
  init(&p);

  mpi_halo_init(&p);
  arrays_allocate(&p);
  init_coeff(&p);
  domain_data_fill(&p);

  TSList[p.target_ts].func(&p);

  fp = fopen(path, "wb");
  if (fp == NULL) { perror("ref_dump"); return 2; }
  hdr[0] = 0x47495249; hdr[1] = (int32_t)sizeof(real_t);
  hdr[2] = p.ldomain_shape[0]; hdr[3] = p.ldomain_shape[1]; hdr[4] = p.ldomain_shape[2];
  hdr[5] = p.stencil.r; hdr[6] = p.nt; hdr[7] = p.target_kernel;
  fwrite(hdr, sizeof(int32_t), 8, fp);
  fwrite(p.U1, sizeof(real_t), p.stencil.type == SOLAR ? p.ln_domain * 24lu : p.ln_domain, fp);
  fclose(fp);
  MPI_Finalize();
  return 0;
}
