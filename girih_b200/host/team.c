/*
 * team.c -- the "ranks" of one mwd_kernel process: one host thread per GPU.  Stands in for the few
 * MPI services the reference's harness uses around the time steppers: MPI_Barrier
 * (src/performance.c:73,77), MPI_Reduce of timers (src/utils.c:844-860), MPI_Bcast, and the gather
 * of sub-domains for verification (src/verification.c:955-1040).  The data path between GPUs is
 * NCCL inside libgirih_cuda, not this file.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "girih_host.h"

static struct {
  int n;
  pthread_barrier_t bar;
  double *slots;   /* n * 16 doubles */
  void *shared;
  unsigned char bcast[4096];
} T = {1, {{0}}, NULL, NULL, {0}};

void team_init(int nranks) {
  T.n = nranks;
  pthread_barrier_init(&T.bar, NULL, (unsigned)nranks);
  T.slots = (double *)calloc((size_t)nranks * 16, sizeof(double));
}

void team_barrier(void) {
  if (T.n > 1) pthread_barrier_wait(&T.bar);
}

void team_reduce(const double *in, double *max, double *min, double *sum, int n, int rank) {
  int i, r;
  if (n > 16) n = 16;
  for (i = 0; i < n; i++) T.slots[rank * 16 + i] = in[i];
  team_barrier();
  for (i = 0; i < n; i++) {
    double mx = T.slots[i], mn = T.slots[i], s = 0;
    for (r = 0; r < T.n; r++) {
      const double v = T.slots[r * 16 + i];
      if (v > mx) mx = v;
      if (v < mn) mn = v;
      s += v;
    }
    if (max) max[i] = mx;
    if (min) min[i] = mn;
    if (sum) sum[i] = s;
  }
  team_barrier();
}

void team_bcast(void *buf, size_t len, int root, int rank) {
  if (T.n == 1) return;
  if (len > sizeof(T.bcast)) {   /* never truncate silently: the receivers would act on a partial message */
    fprintf(stderr, "ERROR: team_bcast of %zu bytes exceeds the %zu-byte slot\n", len, sizeof(T.bcast));
    exit(1);
  }
  if (rank == root) memcpy(T.bcast, buf, len);
  team_barrier();
  if (rank != root) memcpy(buf, T.bcast, len);
  team_barrier();
}

void team_allgather(const void *mine, size_t len, void *all, int rank) {
  static unsigned char slots[64][256];
  int r;
  if (len > 256 || T.n > 64) {   /* the callers size `all` for T.n * len bytes and read all of it */
    fprintf(stderr, "ERROR: team_allgather: %zu bytes x %d ranks exceed the 256-byte x 64-rank slots\n", len, T.n);
    exit(1);
  }
  memcpy(slots[rank], mine, len);
  team_barrier();
  for (r = 0; r < T.n; r++) memcpy((unsigned char *)all + (size_t)r * len, slots[r], len);
  team_barrier();
}

void *team_shared_alloc(size_t bytes, int rank) {
  void *p;
  if (rank == 0) T.shared = malloc(bytes ? bytes : 1);
  team_barrier();
  p = T.shared;
  team_barrier();
  return p;
}

void team_shared_free(void *ptr, int rank) {
  team_barrier();
  if (rank == 0) free(ptr);
}

struct launch { int rank; void (*fn)(int, void *); void *arg; };
static void *trampoline(void *v) {
  struct launch *l = (struct launch *)v;
  l->fn(l->rank, l->arg);
  return NULL;
}

void team_run(int nranks, void (*fn)(int rank, void *arg), void *arg) {
  int r;
  pthread_t *th = (pthread_t *)calloc((size_t)nranks, sizeof(pthread_t));
  struct launch *ls = (struct launch *)calloc((size_t)nranks, sizeof(struct launch));
  team_init(nranks);
  for (r = 1; r < nranks; r++) {
    ls[r].rank = r; ls[r].fn = fn; ls[r].arg = arg;
    pthread_create(&th[r], NULL, trampoline, &ls[r]);
  }
  fn(0, arg);
  for (r = 1; r < nranks; r++) pthread_join(th[r], NULL);
  free(th);
  free(ls);
}
