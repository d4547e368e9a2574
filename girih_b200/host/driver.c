/*
 * driver.c -- main() of mwd_kernel.  Follows src/driver.c:21-82: defaults -> command line ->
 * topology check -> init -> verify or performance_test.  MPI_Init and the MPI ranks become host
 * threads, one per GPU (--npx * --npy * --npz), started after the command line has been parsed once.
 */
#include <stdlib.h>
#include <string.h>

#include "girih_host.h"

static void rank_main(int rank, void *arg) {
  Parameters p = *(const Parameters *)arg;   /* every rank starts from the parsed parameters */
  p.mpi_rank = rank;
  p.mpi_size = p.t.shape[0] * p.t.shape[1] * p.t.shape[2];
  if (rank != 0) p.verbose = 0;
  init(&p);
  if (p.verify != 0) verify(&p);
  else performance_test(&p);
}

int main(int argc, char **argv) {
  Parameters base;
  int nranks;
  memset(&base, 0, sizeof(base));
  base.mpi_size = 1;
  param_default(&base);
  parse_args(argc, argv, &base);   /* --help / --list / bad flags exit here with status 0 */
  nranks = base.t.shape[0] * base.t.shape[1] * base.t.shape[2];
  if (base.t.shape[0] < 1 || base.t.shape[1] < 1 || base.t.shape[2] < 1) nranks = 0;
  if (nranks < 1) {
    fprintf(stderr, "ERROR: requested MPI topology shape does not match the available processes count: \n\tRequested:%03d \n\tAvailable:%03d\n",
            nranks, 1);
    return 1;
  }
  team_run(nranks, rank_main, &base);
  return 0;
}
