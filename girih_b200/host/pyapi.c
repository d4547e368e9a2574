/*
 * pyapi.c -- flat entry points of libgirih_host_{sp,dp}.so for foreign-function callers (the Python
 * package, tests, bench.py): the reference's host-side initialisation (src/utils.c init /
 * init_coeff / domain_data_fill) with plain arguments instead of the Parameters struct.
 */
#include <stdlib.h>
#include <string.h>

#include "girih_host.h"

static void setup_topo(Parameters *p, int kernel, const int gstencil[3], int rank, const int dims[3], int alignment,
                       int padding) {
  memset(p, 0, sizeof(*p));
  p->mpi_rank = rank;
  p->mpi_size = dims[0] * dims[1] * dims[2];
  param_default(p);
  p->verbose = 0;
  p->target_kernel = kernel;
  p->stencil_shape[0] = gstencil[0];
  p->stencil_shape[1] = gstencil[1];
  p->stencil_shape[2] = gstencil[2];
  p->t.shape[0] = dims[0];
  p->t.shape[1] = dims[1];
  p->t.shape[2] = dims[2];
  p->alignment = alignment;
  p->array_padding = padding;
  init(p);
}

static void setup(Parameters *p, int kernel, const int gstencil[3], int rank, int nranks, int alignment, int padding) {
  const int dims[3] = {1, 1, nranks};
  setup_topo(p, kernel, gstencil, rank, dims, alignment, padding);
}

/* the same three services for rank `rank` of an (npx, npy, npz) topology (--npx/--npy/--npz): shapes and global
 * begin, coefficient array size, fill.  coords[3] returns the rank's position (MPI_Cart_coords order). */
int girih_host_shapes_topo(int kernel, const int gstencil[3], int rank, const int dims[3], int alignment, int padding,
                           int lstencil[3], int ldomain[3], int gb[3], int coords[3]) {
  Parameters p;
  int d;
  setup_topo(&p, kernel, gstencil, rank, dims, alignment, padding);
  for (d = 0; d < 3; d++) {
    lstencil[d] = p.lstencil_shape[d]; ldomain[d] = p.ldomain_shape[d]; gb[d] = p.gb[d]; coords[d] = p.t.rank_coords[d];
  }
  return 0;
}

unsigned long long girih_host_coef_size_topo(int kernel, const int gstencil[3], int rank, const int dims[3], int alignment,
                                             int padding) {
  Parameters p;
  setup_topo(&p, kernel, gstencil, rank, dims, alignment, padding);
  return coef_array_size(&p);
}

int girih_host_fill_topo(int kernel, const int gstencil[3], int rank, const int dims[3], int alignment, int padding,
                         void *U1, void *U2, void *U3, void *coef) {
  Parameters p;
  setup_topo(&p, kernel, gstencil, rank, dims, alignment, padding);
  if (p.stencil.time_order == 2 && U3 == NULL) return 1;
  p.U1 = (real_t *)U1; p.U2 = (real_t *)U2; p.U3 = (real_t *)U3; p.coef = (real_t *)coef;
  init_coeff(&p);
  domain_data_fill(&p);
  return 0;
}

int girih_host_elem_size(void) { return (int)sizeof(real_t); }

/* local shapes of z-slab `rank` of `nranks`: lstencil[3], ldomain[3], gb[3] */
int girih_host_shapes(int kernel, const int gstencil[3], int rank, int nranks, int alignment, int padding,
                      int lstencil[3], int ldomain[3], int gb[3]) {
  Parameters p;
  int d;
  setup(&p, kernel, gstencil, rank, nranks, alignment, padding);
  for (d = 0; d < 3; d++) { lstencil[d] = p.lstencil_shape[d]; ldomain[d] = p.ldomain_shape[d]; gb[d] = p.gb[d]; }
  return 0;
}

unsigned long long girih_host_coef_size(int kernel, const int gstencil[3], int rank, int nranks, int alignment, int padding) {
  Parameters p;
  setup(&p, kernel, gstencil, rank, nranks, alignment, padding);
  return coef_array_size(&p);
}

/* init_coeff + domain_data_fill into caller-owned arrays of the shapes reported above */
int girih_host_fill(int kernel, const int gstencil[3], int rank, int nranks, int alignment, int padding,
                    void *U1, void *U2, void *U3, void *coef) {
  Parameters p;
  setup(&p, kernel, gstencil, rank, nranks, alignment, padding);
  if (p.stencil.time_order == 2 && U3 == NULL) return 1;
  p.U1 = (real_t *)U1; p.U2 = (real_t *)U2; p.U3 = (real_t *)U3; p.coef = (real_t *)coef;
  init_coeff(&p);
  domain_data_fill(&p);
  return 0;
}

/* the diamond stepper's rounding of nt, src/kernels/diamond_utils.c:1042-1056 */
int girih_host_diamond_nt(int nt, int t_dim) {
  const int remain = (nt - 2) % ((t_dim + 1) * 2);
  return remain != 0 ? nt + (t_dim + 1) * 2 - remain : nt;
}
