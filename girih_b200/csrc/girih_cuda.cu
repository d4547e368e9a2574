// girih_cuda.cu -- the C ABI declared in include/girih_cuda.h: context, HBM layout, transfers,
// time steppers, z-slab halo exchange.  Kernels live in kernels_*.cuh.
//
// Built only for sm_100a:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include "../../include/girih_cuda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels_naive.cuh"
#include "launch.h"
#include "layout.h"
#include "nccl_dyn.h"
#include "stencil_expr.cuh"

using namespace girih;

// ------------------------------------------------------------------------------------------------
// operator table (stencil_info_list[], src/kernels/stencils.c:260-271)
// ------------------------------------------------------------------------------------------------
static const girih_kernel_desc KERNELS[8] = {
    // name  r to nd shape       coeff                       nca ncs words tfuse gpu
    {"star", 4, 2, 3, GIRIH_STAR, GIRIH_COEF_CONSTANT, 0, 5, 4, 1, 1},
    {"star", 1, 1, 2, GIRIH_STAR, GIRIH_COEF_CONSTANT, 0, 2, 2, 4, 1},
    {"star", 1, 1, 4, GIRIH_STAR, GIRIH_COEF_VARIABLE, 2, 0, 4, 3, 1},
    {"star", 1, 1, 6, GIRIH_STAR, GIRIH_COEF_VARIABLE_AXSYM, 4, 0, 6, 3, 1},
    {"star", 4, 1, 15, GIRIH_STAR, GIRIH_COEF_VARIABLE_AXSYM, 13, 0, 15, 1, 1},
    {"star", 1, 1, 9, GIRIH_STAR, GIRIH_COEF_VARIABLE_NOSYM, 7, 0, 9, 3, 1},
    {"star", 1, 1, 40, GIRIH_STAR, GIRIH_COEF_SOLAR, 0, 0, 104, 1, 1},   // solar: 28 complex arrays, not counted in nca
    {"box", 1, 1, 2, GIRIH_BOX, GIRIH_COEF_CONSTANT, 0, 4, 2, 1, 1},
};

extern "C" int girih_kernel_count(void) { return 8; }
extern "C" int girih_kernel_info(int k, girih_kernel_desc *out) {
  if (k < 0 || k >= 8 || out == nullptr) return GIRIH_ERR_ARG;
  *out = KERNELS[k];
  return GIRIH_OK;
}

extern "C" const char *girih_gpu_strerror(int s) {
  switch (s) {
    case GIRIH_OK: return "success";
    case GIRIH_ERR_ARG: return "invalid argument";
    case GIRIH_ERR_NO_DEVICE: return "no CUDA device available (this library has no CPU fallback)";
    case GIRIH_ERR_CUDA: return "CUDA runtime error";
    case GIRIH_ERR_UNSUPPORTED: return "unsupported configuration for the selected stencil";
    case GIRIH_ERR_NCCL: return "NCCL error";
    case GIRIH_ERR_STATE: return "invalid call order";
    case GIRIH_ERR_FRAME: return "fused stepping requires identical boundary frames in U1 and U2";
    default: return "unknown error";
  }
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct EvPair { cudaEvent_t a, b; };

struct girih_gpu_ctx {
  int device = 0, kernel = 0, es = 8, rank = 0, nranks = 1;
  // process topology (src/mpi_utils.c:63-81): dims = (npx, npy, npz), MPI_Cart_create's row-major rank order
  // (z fastest).  Default: z-slabs only.  Set by girih_gpu_set_topology.
  int dims[3] = {1, 1, 1}, coords[3] = {0, 0, 0};
  void *d_pack = nullptr;             // x/y face staging: [send-, send+, recv-, recv+] x pack_elems
  size_t pack_elems = 0;
  int hshape[3] = {0, 0, 0};   // host array shape
  int st[3] = {0, 0, 0};       // local interior
  girih_kernel_desc kd{};
  DevGrid g{};
  size_t arr_elems = 0;
  int halo_max = 0;            // deepest z halo the allocation supports (planes)
  int nz_min = 0;              // thinnest slab of the run (agreed at comm_init): every rank derives the same schedule
  int opt_halo_group = 0;      // fused passes served by one exchange (0 = choose)
  void *dU[2] = {nullptr, nullptr};   // [0] = U1, [1] = U2
  void *dU3 = nullptr, *dCoef = nullptr;
  double cc[5] = {0, 0, 0, 0, 0};
  cudaStream_t s_comp = nullptr, s_comm = nullptr;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_x = nullptr, ev_y = nullptr;
  std::vector<EvPair> comm_ev;
  size_t comm_ev_used = 0;
  std::vector<EvPair> comp_ev;        // one pair around every sweep of a run: "compute" is their sum, not total - comm
  size_t comp_ev_used = 0;
  bool uploaded = false, frames_equal = true, static_halo_done = false;
  // slot 6 (solar): dU[0] holds the reference's ONE array of 12 complex fields, dCoef its 28 complex coefficient arrays,
  // both in the host layout (no re-pitching); solar_n2 = reals per field = 2 * nnx * nny * nnz
  bool solar = false;
  long long solar_n2 = 0;
  bool frames_dirty = false;   // fields were replaced on the device since frames_equal was evaluated (upload_fields / commit_fields)
  unsigned long long *h_fflag = nullptr;   // page-locked, device-mapped word the frame check writes (no copy engine involved:
                                           // a D2H copy would queue behind the asynchronous download of the previous job)
  // NCCL
  ncclComm_t comm = nullptr;
  // options
  int opt_variant = 0, opt_zchunk = 0, opt_tile = 0, opt_overlap = 0, opt_contract = 0;
  // accounting of the last run
  double ms_compute = 0, ms_comm = 0, ms_total = 0;
  int n_kernels = 0, n_passes = 0, n_steps = 0, tfuse_used = 1;
  int tuned_tile[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // per fusion depth, set by girih_gpu_autotune (0 = built-in default)
  int tuned_tfuse = 0;                            // 0 = built-in default depth
  unsigned long long *d_scan = nullptr;
  void *d_stage = nullptr;            // linear copy of one host array (fast-path transfers)
  // pipelined transfers (girih_gpu_prefetch_fields ... girih_gpu_sync_transfers): staging per array and direction,
  // one stream per copy direction, events that hand the staging buffers back and forth with the compute stream
  void *d_in[2] = {nullptr, nullptr}, *d_out[2] = {nullptr, nullptr};
  bool in_pending[2] = {false, false};
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  // halo push (girih_gpu_peer_export / _peer_attach, option "halo_push"): the z neighbours' arrays and flag words
  // mapped into this process; [0] = lower neighbour, [1] = upper neighbour
  int *d_flags = nullptr;                 // [0] written by the lower neighbour, [1] by the upper one, [2] = wait timed out
  void *peer_U[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [neighbour][array]
  int *peer_flags[2] = {nullptr, nullptr};
  int peer_nz[2] = {0, 0};
  bool peer_ipc[2] = {false, false};      // mapped with cudaIpcOpenMemHandle (to be closed)
  int opt_push = 0;
  int opt_copy = 0;                       // option "halo_copy": overlapped passes move their halos with the copy engines
  int opt_zwave = 0, opt_zwave_block = 0; // options "zwave" (time steps in flight, 0 = choose) / "zwave_block" (planes per launch)
  int push_seq = 0;                       // passes signalled so far (the same number on every rank)
  int push_planes = 0;                    // planes the pass being launched pushes to each neighbour (0 = none)
  cudaEvent_t ev_in_ready = nullptr, ev_in_free = nullptr, ev_out_ready = nullptr, ev_out_free = nullptr;
  // exact-tiled fused sweep (kernels_r1x.cuh): inbound edge slots of the co-resident CTAs (they stay in L2), the running
  // slot tag and the flag a CTA raises when a neighbour tile never delivered
  unsigned char *d_xbuf = nullptr;
  size_t xbuf_bytes = 0;
  unsigned xseq = 0;
  int *d_xerr = nullptr;
  int nsm = 148;
  bool xbuf_used = false;
  long long n_exact = 0, n_fused = 0;   // fused passes on exact tiles / all fused passes, since creation
  char err[512] = "";
};

static int fail(girih_gpu_ctx *c, int status, const char *fmt, ...) {
  if (c) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->err, sizeof(c->err), fmt, ap);
    va_end(ap);
  }
  return status;
}
#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(c, GIRIH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                  __FILE__, __LINE__);                                                        \
  } while (0)
#define NC(call)                                                                              \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess)                                                                    \
      return fail(c, GIRIH_ERR_NCCL, "%s failed: %s (%s:%d)", #call,                          \
                  nccl_dyn()->GetErrorString(r_), __FILE__, __LINE__);                        \
  } while (0)

extern "C" const char *girih_gpu_last_error(girih_gpu_ctx *c) { return c ? c->err : ""; }

// rank of the neighbour one step along dimension d (dir = -1 / +1), or -1 at the domain boundary
// (the non-periodic MPI_Cart_shift of src/mpi_utils.c:78-80)
static int neighbour(const girih_gpu_ctx *c, int d, int dir) {
  int q[3] = {c->coords[0], c->coords[1], c->coords[2]};
  q[d] += dir;
  if (q[d] < 0 || q[d] >= c->dims[d]) return -1;
  return (q[0] * c->dims[1] + q[1]) * c->dims[2] + q[2];
}
static bool xy_decomposed(const girih_gpu_ctx *c) { return c->dims[0] > 1 || c->dims[1] > 1; }

extern "C" int girih_gpu_count(int *n) {
  int k = 0;
  cudaError_t e = cudaGetDeviceCount(&k);
  if (n) *n = (e == cudaSuccess) ? k : 0;
  if (e != cudaSuccess || k == 0) return GIRIH_ERR_NO_DEVICE;
  return GIRIH_OK;
}


extern "C" int girih_gpu_create(girih_gpu_ctx **out, int device, int target_kernel, int elem_size,
                                const int st[3], const int ds[3], int rank, int nranks) {
  if (!out || !st || !ds) return GIRIH_ERR_ARG;
  *out = nullptr;
  if (target_kernel < 0 || target_kernel >= 8) return GIRIH_ERR_ARG;
  if (elem_size != 4 && elem_size != 8) return GIRIH_ERR_ARG;
  if (!KERNELS[target_kernel].gpu_supported) return GIRIH_ERR_UNSUPPORTED;
  if (nranks < 1 || rank < 0 || rank >= nranks) return GIRIH_ERR_ARG;
  const bool solar = KERNELS[target_kernel].coeff == GIRIH_COEF_SOLAR;
  // solar: one rank (the reference's halo exchange moves U1 / U2 of one real per cell, src/mpi_utils.c:173-200, not the
  // 12-field array) and no x padding (src/utils.c:359-361)
  if (solar && nranks != 1) return GIRIH_ERR_UNSUPPORTED;
  if (solar && ds[0] != st[0] + 2) return GIRIH_ERR_ARG;
  const int r = KERNELS[target_kernel].r;
  if (st[0] < 1 || st[1] < 1 || st[2] < 1) return GIRIH_ERR_ARG;
  if (ds[0] < st[0] + 2 * r || ds[1] != st[1] + 2 * r || ds[2] != st[2] + 2 * r) return GIRIH_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return GIRIH_ERR_NO_DEVICE;
  if (device < 0 || device >= ndev) return GIRIH_ERR_ARG;

  girih_gpu_ctx *c = new girih_gpu_ctx();
  c->device = device; c->kernel = target_kernel; c->es = elem_size; c->rank = rank; c->nranks = nranks;
  c->dims[2] = nranks; c->coords[2] = rank;
  c->kd = KERNELS[target_kernel];
  for (int d = 0; d < 3; ++d) { c->hshape[d] = ds[d]; c->st[d] = st[d]; }
  auto bail = [&](int status) { girih_gpu_destroy(c); return status; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(GIRIH_ERR_CUDA);

  // HBM layout (see DevGrid and layout.h)
  const int guard = c->kd.max_tfuse * r;      // deepest halo / overlap any stepper uses
  DevGrid &g = c->g;
  const int zguard = make_dev_grid(g, st, r, c->kd.max_tfuse, elem_size, rank, nranks);
  c->halo_max = (nranks > 1) ? std::min(zguard, std::max(g.nz, std::max(guard, r))) : zguard;
  c->nz_min = g.nz;
  c->arr_elems = (size_t)g.pxy * g.nz_dev;
  if (nranks > 1 && g.nz < std::max(guard, r)) {
    fail(c, GIRIH_ERR_ARG, "slab of %d planes is thinner than the deepest halo (%d)", g.nz, std::max(guard, r));
    return bail(GIRIH_ERR_ARG);
  }

  const size_t bytes = c->arr_elems * elem_size;
  cudaError_t e = cudaSuccess;
  c->solar = solar;
  if (solar) {
    c->solar_n2 = 2LL * ds[0] * ds[1] * ds[2];
    e = cudaMalloc(&c->dU[0], (size_t)c->solar_n2 * 12 * elem_size);
    if (e == cudaSuccess) e = cudaMalloc(&c->dCoef, (size_t)c->solar_n2 * 28 * elem_size);
  }
  for (int i = 0; i < 2 && e == cudaSuccess && !solar; ++i) {
    e = cudaMalloc(&c->dU[i], bytes);
    if (e == cudaSuccess) e = cudaMemset(c->dU[i], 0, bytes);
  }
  if (e == cudaSuccess && c->kd.time_order == 2) {
    e = cudaMalloc(&c->dU3, bytes);
    if (e == cudaSuccess) e = cudaMemset(c->dU3, 0, bytes);
  }
  if (e == cudaSuccess && c->kd.n_coef_arrays > 0) {
    e = cudaMalloc(&c->dCoef, bytes * c->kd.n_coef_arrays);
    if (e == cudaSuccess) e = cudaMemset(c->dCoef, 0, bytes * c->kd.n_coef_arrays);
  }
  if (e == cudaSuccess) e = cudaMalloc(&c->d_scan, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking);
  if (e == cudaSuccess) {   // the exchange stream outranks the sweep: its few CTAs go first when SM slots free up
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    e = cudaStreamCreateWithPriority(&c->s_comm, cudaStreamNonBlocking, hi);
  }
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev_t0);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev_t1);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_y, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();   // memsets above are asynchronous to the host
  if (e != cudaSuccess) {
    fprintf(stderr, "girih_gpu_create: %s\n", cudaGetErrorString(e));
    return bail(GIRIH_ERR_CUDA);
  }
  *out = c;
  return GIRIH_OK;
}

extern "C" void girih_gpu_destroy(girih_gpu_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->s_comp) cudaStreamSynchronize(c->s_comp);
  if (c->s_comm) cudaStreamSynchronize(c->s_comm);
  if (c->comm && nccl_dyn()) nccl_dyn()->CommDestroy(c->comm);
  for (int i = 0; i < 2; ++i) if (c->dU[i]) cudaFree(c->dU[i]);
  if (c->dU3) cudaFree(c->dU3);
  if (c->dCoef) cudaFree(c->dCoef);
  if (c->d_scan) cudaFree(c->d_scan);
  if (c->h_fflag) cudaFreeHost(c->h_fflag);
  if (c->d_xbuf) cudaFree(c->d_xbuf);
  if (c->d_xerr) cudaFree(c->d_xerr);
  if (c->d_stage) cudaFree(c->d_stage);
  if (c->d_pack) cudaFree(c->d_pack);
  if (c->s_h2d) cudaStreamSynchronize(c->s_h2d);
  if (c->s_d2h) cudaStreamSynchronize(c->s_d2h);
  for (int i = 0; i < 2; ++i) {
    if (c->d_in[i]) cudaFree(c->d_in[i]);
    if (c->d_out[i]) cudaFree(c->d_out[i]);
  }
  for (cudaEvent_t ev : {c->ev_in_ready, c->ev_in_free, c->ev_out_ready, c->ev_out_free})
    if (ev) cudaEventDestroy(ev);
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  girih_gpu_peer_detach(c);
  if (c->d_flags) cudaFree(c->d_flags);
  for (auto &p : c->comm_ev) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto &p : c->comp_ev) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->ev_x) cudaEventDestroy(c->ev_x);
  if (c->ev_y) cudaEventDestroy(c->ev_y);
  if (c->s_comp) cudaStreamDestroy(c->s_comp);
  if (c->s_comm) cudaStreamDestroy(c->s_comm);
  delete c;
}

// Process topology of the run (the reference's --npx/--npy/--npz, src/mpi_utils.c:63-81).  Without this call the
// ranks form z-slabs.  With npx or npy > 1 the single-step steppers exchange r-deep x and y faces as well
// (src/mpi_utils.c:116-170) and temporal fusion is not offered, like the reference, whose diamond stepper
// does not decompose x either (src/kernels/diamond_utils.c:1035-1040).
extern "C" int girih_gpu_set_topology(girih_gpu_ctx *c, const int dims[3], const int coords[3]) {
  if (!c || !dims || !coords) return GIRIH_ERR_ARG;
  if (c->comm) return fail(c, GIRIH_ERR_STATE, "set_topology after comm_init");
  for (int d = 0; d < 3; ++d)
    if (dims[d] < 1 || coords[d] < 0 || coords[d] >= dims[d]) return fail(c, GIRIH_ERR_ARG, "bad topology");
  if (dims[0] * dims[1] * dims[2] != c->nranks) return fail(c, GIRIH_ERR_ARG, "topology does not match the rank count");
  if ((coords[0] * dims[1] + coords[1]) * dims[2] + coords[2] != c->rank)
    return fail(c, GIRIH_ERR_ARG, "coordinates do not match the rank (row-major, z fastest)");
  for (int d = 0; d < 3; ++d) { c->dims[d] = dims[d]; c->coords[d] = coords[d]; }
  DevGrid &g = c->g;
  g.zlo = (coords[2] == 0) ? g.Z0 : -(1 << 30);
  g.zhi = (coords[2] == dims[2] - 1) ? g.Z0 + g.nz : (1 << 30);
  return GIRIH_OK;
}

extern "C" int girih_gpu_set_option(girih_gpu_ctx *c, const char *key, int value) {
  if (!c || !key) return GIRIH_ERR_ARG;
  if (!strcmp(key, "variant")) c->opt_variant = value;
  else if (!strcmp(key, "zchunk")) c->opt_zchunk = value;
  else if (!strcmp(key, "tile")) c->opt_tile = value;
  else if (!strcmp(key, "overlap")) c->opt_overlap = value;
  else if (!strcmp(key, "contract")) c->opt_contract = (value != 0);
  else if (!strcmp(key, "halo_group")) c->opt_halo_group = value;
  else if (!strcmp(key, "halo_push")) c->opt_push = value;
  else if (!strcmp(key, "halo_copy")) c->opt_copy = value;
  else if (!strcmp(key, "zwave")) c->opt_zwave = value;
  else if (!strcmp(key, "zwave_block")) c->opt_zwave_block = value;
  else return fail(c, GIRIH_ERR_ARG, "unknown option '%s'", key);
  return GIRIH_OK;
}

// ------------------------------------------------------------------------------------------------
// transfers
// ------------------------------------------------------------------------------------------------
// host array (reference layout, hshape) <-> device array (DevGrid layout); all planes of the host
// array are moved, i.e. the interior plus the r-deep frame / inter-slab halo.
static cudaError_t copy3d(girih_gpu_ctx *c, void *dev, const void *host_c, bool to_device,
                          cudaStream_t s, bool async) {
  const DevGrid &g = c->g;
  const int r = g.r, es = c->es;
  void *host = const_cast<void *>(host_c);
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  cudaPitchedPtr hp = make_cudaPitchedPtr(host, (size_t)c->hshape[0] * es, (size_t)c->hshape[0] * es, c->hshape[1]);
  cudaPitchedPtr dp = make_cudaPitchedPtr(dev, (size_t)g.px * es, (size_t)g.px * es, g.ny_dev);
  cudaPos hpos = make_cudaPos(0, 0, 0);
  cudaPos dpos = make_cudaPos((size_t)(g.X0 - r) * es, g.Y0 - r, g.Z0 - r);
  p.extent = make_cudaExtent((size_t)(g.nx + 2 * r) * es, c->hshape[1], c->hshape[2]);
  if (to_device) { p.srcPtr = hp; p.srcPos = hpos; p.dstPtr = dp; p.dstPos = dpos; p.kind = cudaMemcpyHostToDevice; }
  else           { p.srcPtr = dp; p.srcPos = dpos; p.dstPtr = hp; p.dstPos = hpos; p.kind = cudaMemcpyDeviceToHost; }
  return async ? cudaMemcpy3DAsync(&p, s) : cudaMemcpy3D(&p);
}

template <typename R>
static bool frames_match(const girih_gpu_ctx *c, const R *a, const R *b) {
  // true iff U1 and U2 agree on every cell the steppers never write: the Dirichlet frame
  // (not the halo planes between slabs, which are interior points of the global domain)
  const int nnx = c->hshape[0], nny = c->hshape[1], nnz = c->hshape[2], r = c->g.r;
  const int nx = c->st[0];
  const bool zfirst = c->coords[2] == 0, zlast = c->coords[2] == c->dims[2] - 1;
  for (int k = 0; k < nnz; ++k) {
    const bool kframe = (zfirst && k < r) || (zlast && k >= nnz - r);
    for (int j = 0; j < nny; ++j) {
      const bool jframe = (j < r) || (j >= nny - r);
      const size_t row = ((size_t)k * nny + j) * nnx;
      if (kframe || jframe) {
        if (memcmp(a + row, b + row, sizeof(R) * (size_t)(nx + 2 * r)) != 0) return false;
      } else {
        if (memcmp(a + row, b + row, sizeof(R) * r) != 0) return false;
        if (memcmp(a + row + nx + r, b + row + nx + r, sizeof(R) * r) != 0) return false;
      }
    }
  }
  return true;
}

extern "C" int girih_gpu_upload(girih_gpu_ctx *c, const void *U1, const void *U2, const void *U3,
                                const void *coef) {
  if (!c || !U1) return GIRIH_ERR_ARG;
  if (c->solar) {   // one field array (U2 == 0 in the reference, src/utils.c:171) and the coefficient arrays, layout unchanged
    if (!coef) return fail(c, GIRIH_ERR_ARG, "coef is required");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->dU[0], U1, (size_t)c->solar_n2 * 12 * c->es, cudaMemcpyHostToDevice, c->s_comp));
    CU(cudaMemcpyAsync(c->dCoef, coef, (size_t)c->solar_n2 * 28 * c->es, cudaMemcpyHostToDevice, c->s_comp));
    CU(cudaStreamSynchronize(c->s_comp));
    c->uploaded = true;
    return GIRIH_OK;
  }
  if (!U2) return GIRIH_ERR_ARG;
  if (c->kd.time_order == 2 && !U3) return fail(c, GIRIH_ERR_ARG, "U3 (roc2) is required for time_order 2");
  if ((c->kd.n_coef_arrays > 0 || c->kd.n_coef_scalars > 0) && !coef) return fail(c, GIRIH_ERR_ARG, "coef is required");
  CU(cudaSetDevice(c->device));
  // All transfers go through the context's own (non-blocking) stream: a synchronous cudaMemcpy from
  // pageable memory may return while the DMA is still in flight, and kernels on a non-blocking stream
  // are not ordered behind the legacy default stream.
  CU(copy3d(c, c->dU[0], U1, true, c->s_comp, true));
  CU(copy3d(c, c->dU[1], U2, true, c->s_comp, true));
  if (c->kd.time_order == 2) CU(copy3d(c, c->dU3, U3, true, c->s_comp, true));
  const size_t ln = (size_t)c->hshape[0] * c->hshape[1] * c->hshape[2];
  for (int m = 0; m < c->kd.n_coef_arrays; ++m)
    CU(copy3d(c, (char *)c->dCoef + (size_t)m * c->arr_elems * c->es,
              (const char *)coef + (size_t)m * ln * c->es, true, c->s_comp, true));
  CU(cudaStreamSynchronize(c->s_comp));
  for (int m = 0; m < c->kd.n_coef_scalars; ++m)
    c->cc[m] = (c->es == 8) ? ((const double *)coef)[m] : (double)((const float *)coef)[m];
  c->frames_equal = (c->es == 8) ? frames_match(c, (const double *)U1, (const double *)U2)
                                 : frames_match(c, (const float *)U1, (const float *)U2);
  c->frames_dirty = false;
  c->uploaded = true;
  c->static_halo_done = false;
  return GIRIH_OK;
}

// ------------------------------------------------------------------------------------------------
// fast transfers for page-locked host arrays: ONE linear DMA of the whole host array (full PCIe rate,
// a pitched 3-D copy moves row by row) plus a device kernel that converts between the reference's
// host layout and the DevGrid layout.  Used by girih_gpu_upload_fields / girih_gpu_download.
// ------------------------------------------------------------------------------------------------
template <typename R, bool TO_DEVICE>
__global__ void k_repitch(DevGrid g, R *__restrict__ dev, R *__restrict__ lin, int hx, int hy, int hz) {
  // lin: [hz][hy][hx] (host layout, hx includes the x padding); only the first nx+2r columns are moved
  const int w = g.nx + 2 * g.r;
  const long long rows = (long long)hy * hz;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int y = (int)(row % hy), z = (int)(row / hy);
    R *d = dev + ((long long)(z + g.Z0 - g.r) * g.ny_dev + (y + g.Y0 - g.r)) * g.px + (g.X0 - g.r);
    R *l = lin + row * hx;
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
      if (TO_DEVICE) d[x] = l[x];
      else l[x] = d[x];
    }
  }
}

static int ensure_stage(girih_gpu_ctx *c) {
  if (!c->d_stage) {
    const size_t bytes = (size_t)c->hshape[0] * c->hshape[1] * c->hshape[2] * c->es;
    CU(cudaMalloc(&c->d_stage, bytes));
    // the x padding columns of the host layout are never touched by the steppers; they travel through this
    // buffer unchanged (zero until an upload brings the host's own padding, which GIRIH's fill leaves zero)
    CU(cudaMemsetAsync(c->d_stage, 0, bytes, c->s_comp));
  }
  return GIRIH_OK;
}

static int fast_copy(girih_gpu_ctx *c, void *dev, void *host, bool to_device) {
  int rc = ensure_stage(c);
  if (rc) return rc;
  const size_t bytes = (size_t)c->hshape[0] * c->hshape[1] * c->hshape[2] * c->es;
  const int grid = 148 * 16;
  const int hx = c->hshape[0], hy = c->hshape[1], hz = c->hshape[2];
  auto repitch = [&](auto kd, auto kf) {   // kd / kf: the double / float instantiation of one direction
    if (c->es == 8) GIRIH_LAUNCH(kd, grid, 256, 0, c->s_comp, c->g, (double *)dev, (double *)c->d_stage, hx, hy, hz);
    else GIRIH_LAUNCH(kf, grid, 256, 0, c->s_comp, c->g, (float *)dev, (float *)c->d_stage, hx, hy, hz);
  };
  if (to_device) {
    CU(cudaMemcpyAsync(c->d_stage, host, bytes, cudaMemcpyHostToDevice, c->s_comp));
    repitch(k_repitch<double, true>, k_repitch<float, true>);
    CU(cudaGetLastError());
  } else {
    repitch(k_repitch<double, false>, k_repitch<float, false>);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->s_comp));
  }
  return GIRIH_OK;
}

extern "C" int girih_gpu_upload_fields(girih_gpu_ctx *c, const void *U1, const void *U2) {
  if (!c || !c->uploaded) return fail(c, GIRIH_ERR_STATE, "upload_fields before upload");
  CU(cudaSetDevice(c->device));
  if (c->solar) {
    if (U1) CU(cudaMemcpyAsync(c->dU[0], U1, (size_t)c->solar_n2 * 12 * c->es, cudaMemcpyHostToDevice, c->s_comp));
    CU(cudaStreamSynchronize(c->s_comp));
    return GIRIH_OK;
  }
  int rc;
  if (U1 && (rc = fast_copy(c, c->dU[0], const_cast<void *>(U1), true))) return rc;
  if (U2 && (rc = fast_copy(c, c->dU[1], const_cast<void *>(U2), true))) return rc;
  CU(cudaStreamSynchronize(c->s_comp));
  if (U1 || U2) c->frames_dirty = true;   // the frames travelled too: re-evaluated on the device before the next fused run
  return GIRIH_OK;
}

// Do U1 and U2 agree on the Dirichlet frame?  Device-side twin of frames_match() for fields that were replaced
// on the device (upload_fields, commit_fields): one flag word, read back only when a fused run needs the answer.
__device__ __forceinline__ bool same_bits(double x, double y) {
  return *reinterpret_cast<const unsigned long long *>(&x) == *reinterpret_cast<const unsigned long long *>(&y);
}
__device__ __forceinline__ bool same_bits(float x, float y) {
  return *reinterpret_cast<const unsigned *>(&x) == *reinterpret_cast<const unsigned *>(&y);
}
template <typename R>
__global__ void k_frame_diff(DevGrid g, const R *__restrict__ a, const R *__restrict__ b, int hy, int hz, int zfirst,
                             int zlast, unsigned long long *flag) {
  const int r = g.r, w = g.nx + 2 * r;
  const long long rows = (long long)hy * hz;
  bool diff = false;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int y = (int)(row % hy), z = (int)(row / hy);
    const bool whole = (zfirst && z < r) || (zlast && z >= hz - r) || (y < r) || (y >= hy - r);
    const long long o = ((long long)(z + g.Z0 - r) * g.ny_dev + (y + g.Y0 - r)) * g.px + (g.X0 - r);
    if (whole) {
      for (int x = threadIdx.x; x < w; x += blockDim.x) diff |= !same_bits(a[o + x], b[o + x]);
    } else if ((int)threadIdx.x < 2 * r) {
      const int x = (int)threadIdx.x < r ? (int)threadIdx.x : w - 2 * r + (int)threadIdx.x;
      diff |= !same_bits(a[o + x], b[o + x]);
    }
  }
  if (diff) *flag = 1ull;
}

static int refresh_frames_equal(girih_gpu_ctx *c) {
  if (!c->frames_dirty) return GIRIH_OK;
  if (!c->h_fflag) CU(cudaHostAlloc((void **)&c->h_fflag, sizeof(unsigned long long), cudaHostAllocMapped));
  unsigned long long *flag = c->h_fflag;   // unified addressing: the host pointer is valid on the device
  *flag = 0ull;
  const int zf = c->coords[2] == 0, zl = c->coords[2] == c->dims[2] - 1;
  if (c->es == 8) {
    auto k = k_frame_diff<double>;
    GIRIH_LAUNCH(k, 148 * 8, 128, 0, c->s_comp, c->g, (const double *)c->dU[0], (const double *)c->dU[1], c->hshape[1], c->hshape[2], zf, zl, flag);
  } else {
    auto k = k_frame_diff<float>;
    GIRIH_LAUNCH(k, 148 * 8, 128, 0, c->s_comp, c->g, (const float *)c->dU[0], (const float *)c->dU[1], c->hshape[1], c->hshape[2], zf, zl, flag);
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->s_comp));
  c->frames_equal = (*(volatile unsigned long long *)flag == 0);
  c->frames_dirty = false;
  return GIRIH_OK;
}

extern "C" int girih_gpu_download(girih_gpu_ctx *c, void *U1, void *U2) {
  if (!c) return GIRIH_ERR_ARG;
  if (!c->uploaded) return fail(c, GIRIH_ERR_STATE, "download before upload");
  CU(cudaSetDevice(c->device));
  if (c->solar) {
    if (U1) CU(cudaMemcpyAsync(U1, c->dU[0], (size_t)c->solar_n2 * 12 * c->es, cudaMemcpyDeviceToHost, c->s_comp));
    CU(cudaStreamSynchronize(c->s_comp));
    return GIRIH_OK;
  }
  int rc;
  if (U1 && (rc = fast_copy(c, c->dU[0], U1, false))) return rc;
  if (U2 && (rc = fast_copy(c, c->dU[1], U2, false))) return rc;
  CU(cudaStreamSynchronize(c->s_comp));
  return GIRIH_OK;
}

// ------------------------------------------------------------------------------------------------
// Pipelined transfers for a stream of independent jobs on one context.  The copies of job i+1 (host -> device)
// and of job i-1 (device -> host) run on their own streams, i.e. on the two copy engines, underneath the
// sweeps of job i; the layout conversions stay on the compute stream (0.4 ms per array at 512^3).
//   prefetch_fields  asynchronous DMA of the next job's fields into staging (page-locked host arrays, valid
//                    until commit_fields returns)
//   commit_fields    the prefetched fields become the device arrays (ordered behind the DMA; the staging
//                    buffers are handed back to the next prefetch by an event)
//   download_async   layout conversion on the compute stream, then asynchronous DMA to the host
//   sync_transfers   blocks until every transfer issued so far has completed
// ------------------------------------------------------------------------------------------------
static int ensure_pipeline(girih_gpu_ctx *c) {
  if (c->s_h2d) return GIRIH_OK;
  CU(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->ev_in_ready, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_in_free, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_out_ready, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_out_free, cudaEventDisableTiming));
  return GIRIH_OK;
}

static cudaError_t launch_repitch(girih_gpu_ctx *c, void *dev, void *lin, bool to_device) {
  const int grid = 148 * 16, hx = c->hshape[0], hy = c->hshape[1], hz = c->hshape[2];
  if (c->es == 8) {
    if (to_device) { auto k = k_repitch<double, true>; GIRIH_LAUNCH(k, grid, 256, 0, c->s_comp, c->g, (double *)dev, (double *)lin, hx, hy, hz); }
    else { auto k = k_repitch<double, false>; GIRIH_LAUNCH(k, grid, 256, 0, c->s_comp, c->g, (double *)dev, (double *)lin, hx, hy, hz); }
  } else {
    if (to_device) { auto k = k_repitch<float, true>; GIRIH_LAUNCH(k, grid, 256, 0, c->s_comp, c->g, (float *)dev, (float *)lin, hx, hy, hz); }
    else { auto k = k_repitch<float, false>; GIRIH_LAUNCH(k, grid, 256, 0, c->s_comp, c->g, (float *)dev, (float *)lin, hx, hy, hz); }
  }
  return cudaGetLastError();
}

extern "C" int girih_gpu_prefetch_fields(girih_gpu_ctx *c, const void *U1, const void *U2) {
  if (c && c->solar) return fail(c, GIRIH_ERR_UNSUPPORTED, "pipelined transfers are not offered for the solar slot");
  if (!c || !c->uploaded) return fail(c, GIRIH_ERR_STATE, "prefetch_fields before upload");
  CU(cudaSetDevice(c->device));
  int rc = ensure_pipeline(c);
  if (rc) return rc;
  const size_t bytes = (size_t)c->hshape[0] * c->hshape[1] * c->hshape[2] * c->es;
  const void *host[2] = {U1, U2};
  CU(cudaStreamWaitEvent(c->s_h2d, c->ev_in_free, 0));   // the previous commit has drained the staging buffers
  for (int i = 0; i < 2; ++i) {
    c->in_pending[i] = host[i] != nullptr;
    if (!host[i]) continue;
    if (!c->d_in[i]) CU(cudaMalloc(&c->d_in[i], bytes));
    CU(cudaMemcpyAsync(c->d_in[i], host[i], bytes, cudaMemcpyHostToDevice, c->s_h2d));
  }
  CU(cudaEventRecord(c->ev_in_ready, c->s_h2d));
  return GIRIH_OK;
}

extern "C" int girih_gpu_commit_fields(girih_gpu_ctx *c) {
  if (c && c->solar) return fail(c, GIRIH_ERR_UNSUPPORTED, "pipelined transfers are not offered for the solar slot");
  if (!c || !c->uploaded) return fail(c, GIRIH_ERR_STATE, "commit_fields before upload");
  if (!c->s_h2d) return fail(c, GIRIH_ERR_STATE, "commit_fields without prefetch_fields");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamWaitEvent(c->s_comp, c->ev_in_ready, 0));
  for (int i = 0; i < 2; ++i) {
    if (!c->in_pending[i]) continue;
    CU(launch_repitch(c, c->dU[i], c->d_in[i], true));
    c->in_pending[i] = false;
    c->frames_dirty = true;
  }
  CU(cudaEventRecord(c->ev_in_free, c->s_comp));
  return GIRIH_OK;
}

extern "C" int girih_gpu_download_async(girih_gpu_ctx *c, void *U1, void *U2) {
  if (c && c->solar) return fail(c, GIRIH_ERR_UNSUPPORTED, "pipelined transfers are not offered for the solar slot");
  if (!c || !c->uploaded) return fail(c, GIRIH_ERR_STATE, "download before upload");
  CU(cudaSetDevice(c->device));
  int rc = ensure_pipeline(c);
  if (rc) return rc;
  const size_t bytes = (size_t)c->hshape[0] * c->hshape[1] * c->hshape[2] * c->es;
  void *host[2] = {U1, U2};
  CU(cudaStreamWaitEvent(c->s_comp, c->ev_out_free, 0));   // the previous download has left the staging buffers
  for (int i = 0; i < 2; ++i) {
    if (!host[i]) continue;
    if (!c->d_out[i]) {
      CU(cudaMalloc(&c->d_out[i], bytes));
      CU(cudaMemsetAsync(c->d_out[i], 0, bytes, c->s_comp));   // the x padding columns travel as zeros
    }
    CU(launch_repitch(c, c->dU[i], c->d_out[i], false));
  }
  CU(cudaEventRecord(c->ev_out_ready, c->s_comp));
  CU(cudaStreamWaitEvent(c->s_d2h, c->ev_out_ready, 0));
  for (int i = 0; i < 2; ++i)
    if (host[i]) CU(cudaMemcpyAsync(host[i], c->d_out[i], bytes, cudaMemcpyDeviceToHost, c->s_d2h));
  CU(cudaEventRecord(c->ev_out_free, c->s_d2h));
  return GIRIH_OK;
}

extern "C" int girih_gpu_sync_transfers(girih_gpu_ctx *c) {
  if (!c) return GIRIH_ERR_ARG;
  CU(cudaSetDevice(c->device));
  if (c->s_h2d) CU(cudaStreamSynchronize(c->s_h2d));
  CU(cudaStreamSynchronize(c->s_comp));
  if (c->s_d2h) CU(cudaStreamSynchronize(c->s_d2h));
  return GIRIH_OK;
}

// ------------------------------------------------------------------------------------------------
// Halo push over peer memory: compute and exchange in ONE kernel.  Each rank maps its z neighbours' field arrays
// (same process: peer access; other process: CUDA IPC) and the fused sweep stores the boundary planes of every
// pass straight into the neighbours' halo planes over NVLink while it computes (kernels_r1.cuh, R1_PUSH).  What is
// left of the exchange is ordering: a pass may start when both neighbours have finished the previous one (their
// stores into my halos are complete, and they no longer read the halos my pass overwrites).  That is one flag
// word per neighbour in device memory, written by a one-thread kernel behind each pass and awaited by a one-thread
// kernel in front of the next -- no host synchronisation, no NCCL kernel competing for SMs.
//   girih_gpu_peer_export   this rank's handles (GIRIH_PEER_BLOB_BYTES), to be passed to both z neighbours
//   girih_gpu_peer_attach   maps one neighbour (which = 0 lower, 1 upper) from its blob
//   option "halo_push" = 1  fused passes of slot 1 use it (default tiles; needs both attach calls where a
//                           neighbour exists, and comm_init for the first and last exchange of a run)
// ------------------------------------------------------------------------------------------------
struct PeerBlob {
  int magic, pid, device, nz;
  unsigned long long ptr[3];        // dU[0], dU[1], flags (valid inside the exporting process)
  cudaIpcMemHandle_t h[3];
};
static_assert(sizeof(PeerBlob) <= GIRIH_PEER_BLOB_BYTES, "blob size");

// How long a one-thread wait kernel polls its neighbours' flags before it gives up and the run reports an error instead of
// hanging: 60 s of device clock by default (a neighbour may be busy with a first-launch module load, an upload or the tuner),
// GIRIH_FLAG_TIMEOUT_S overrides it (the fixed ~10 s limit could fire spuriously).
static long long flag_wait_cycles() {
  static long long cycles = 0;
  if (cycles == 0) {
    const char *e = getenv("GIRIH_FLAG_TIMEOUT_S");
    double sec = e ? atof(e) : 60.0;
    if (!(sec > 0.0)) sec = 60.0;
    cycles = (long long)(sec * 2.0e9);
  }
  return cycles;
}

__global__ void k_flag_signal(volatile int *dn_slot, volatile int *up_slot, int v) {
  __threadfence_system();   // the sweep's stores into peer memory are ordered before the flag
  if (dn_slot) *dn_slot = v;
  if (up_slot) *up_slot = v;
}
__global__ void k_flag_wait(volatile int *flags, int need_dn, int need_up, int v, long long max_cycles) {
  const long long t0 = clock64();
  while ((need_dn && flags[0] - v < 0) || (need_up && flags[1] - v < 0)) {
    if (clock64() - t0 > max_cycles) { flags[2] = 1; break; }   // a neighbour is gone: give up, the host reports it
  }
  __threadfence_system();
}

extern "C" int girih_gpu_peer_export(girih_gpu_ctx *c, void *blob, size_t len) {
  if (!c || !blob || len < GIRIH_PEER_BLOB_BYTES) return GIRIH_ERR_ARG;
  CU(cudaSetDevice(c->device));
  if (!c->d_flags) {
    CU(cudaMalloc(&c->d_flags, 4 * sizeof(int)));
    CU(cudaMemset(c->d_flags, 0, 4 * sizeof(int)));
  }
  PeerBlob b;
  memset(&b, 0, sizeof(b));
  b.magic = 0x47504252; b.pid = (int)getpid(); b.device = c->device; b.nz = c->g.nz;
  void *ptrs[3] = {c->dU[0], c->dU[1], c->d_flags};
  for (int i = 0; i < 3; ++i) {
    b.ptr[i] = (unsigned long long)(uintptr_t)ptrs[i];
    // the IPC handle only matters to importers in OTHER processes (rank threads of one process use the pointer): where
    // the platform refuses to make one, the blob still serves same-process peers and a foreign importer fails in attach
    if (cudaIpcGetMemHandle(&b.h[i], ptrs[i]) != cudaSuccess) {
      (void)cudaGetLastError();
      memset(&b.h[i], 0, sizeof(b.h[i]));
    }
  }
  memset(blob, 0, len);
  memcpy(blob, &b, sizeof(b));
  return GIRIH_OK;
}

extern "C" int girih_gpu_peer_attach(girih_gpu_ctx *c, int which, const void *blob, size_t len) {
  if (!c || !blob || len < sizeof(PeerBlob) || which < 0 || which > 1) return GIRIH_ERR_ARG;
  PeerBlob b;
  memcpy(&b, blob, sizeof(b));
  if (b.magic != 0x47504252) return fail(c, GIRIH_ERR_ARG, "not a peer blob");
  CU(cudaSetDevice(c->device));
  void *ptrs[3] = {nullptr, nullptr, nullptr};
  if (b.pid == (int)getpid()) {   // rank threads of one process: the pointers are valid here, the devices need peer access
    if (b.device != c->device) {
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, c->device, b.device));
      if (!can) return fail(c, GIRIH_ERR_UNSUPPORTED, "device %d cannot access device %d", c->device, b.device);
      cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
      (void)cudaGetLastError();
    }
    for (int i = 0; i < 3; ++i) ptrs[i] = (void *)(uintptr_t)b.ptr[i];
  } else {
    for (int i = 0; i < 3; ++i) CU(cudaIpcOpenMemHandle(&ptrs[i], b.h[i], cudaIpcMemLazyEnablePeerAccess));
    c->peer_ipc[which] = true;
  }
  c->peer_U[which][0] = ptrs[0];
  c->peer_U[which][1] = ptrs[1];
  c->peer_flags[which] = (int *)ptrs[2];
  c->peer_nz[which] = b.nz;
  return GIRIH_OK;
}

// Unmaps the neighbours (IPC handles are closed).  Importers should detach before an exporter is destroyed: call this
// on every rank, synchronise the ranks, then destroy.  girih_gpu_destroy calls it as well.
extern "C" int girih_gpu_peer_detach(girih_gpu_ctx *c) {
  if (!c) return GIRIH_ERR_ARG;
  cudaSetDevice(c->device);
  if (c->s_comp) cudaStreamSynchronize(c->s_comp);
  for (int n = 0; n < 2; ++n) {
    if (c->peer_ipc[n]) {
      for (int a = 0; a < 2; ++a) if (c->peer_U[n][a]) cudaIpcCloseMemHandle(c->peer_U[n][a]);
      if (c->peer_flags[n]) cudaIpcCloseMemHandle(c->peer_flags[n]);
    }
    c->peer_ipc[n] = false;
    c->peer_U[n][0] = c->peer_U[n][1] = nullptr;
    c->peer_flags[n] = nullptr;
  }
  return GIRIH_OK;
}

// may this run push its halos?  Same answer on every rank: options and topology are set alike by the host
static bool push_enabled(const girih_gpu_ctx *c, int Tmax) {
  if (!c->opt_push || c->nranks == 1 || xy_decomposed(c) || c->kernel != 1 || c->opt_variant == 1 || c->opt_tile != 0) return false;
  if (Tmax * c->g.r > c->nz_min) return false;
  if (Tmax <= 1) return false;   // single-step runs (ts 0/1) exchange through NCCL: nothing is fused, nothing to hide
  const bool need_dn = neighbour(c, 2, -1) >= 0, need_up = neighbour(c, 2, +1) >= 0;
  return c->d_flags && (!need_dn || c->peer_flags[0]) && (!need_up || c->peer_flags[1]);
}

// Halo copy: the overlapped ("halo-first") schedule with the exchange done by the COPY ENGINES instead of NCCL's
// SM kernels.  After the sweep of the two outer parts of the slab, the planes the neighbours read in the next pass are
// copied straight into their halo planes (peer memory mapped by girih_gpu_peer_attach) on the comm stream, followed by a
// one-thread kernel that raises the neighbours' flag words; the sweep of the inner part runs meanwhile on every SM (an
// NCCL kernel in that place takes SMs from a sweep that fills whole waves of 148 and pushes its tail into an extra wave).
// The next pass starts its outer parts once both neighbours' flags have arrived.  Ordering argument:
//   RAW  the neighbour's next outer sweep waits for my flag, raised behind my copy in stream order
//   WAR  my copy overwrites halo planes the neighbour last read in its previous outer sweep, which finished before the
//        neighbour signalled the data I waited for; my own next-but-one outer sweep overwrites the planes this copy
//        reads and therefore waits for the copy's event
static bool copy_enabled(const girih_gpu_ctx *c) {
  if (!c->opt_copy || c->nranks == 1 || xy_decomposed(c)) return false;
  const bool need_dn = neighbour(c, 2, -1) >= 0, need_up = neighbour(c, 2, +1) >= 0;
  return c->d_flags && (!need_dn || c->peer_flags[0]) && (!need_up || c->peer_flags[1]);
}

// ------------------------------------------------------------------------------------------------
// NCCL bootstrap and halo exchange
// ------------------------------------------------------------------------------------------------
extern "C" int girih_gpu_comm_unique_id(void *id, size_t len) {
  if (!id || len < GIRIH_COMM_ID_BYTES) return GIRIH_ERR_ARG;
  if (!nccl_dyn()) return GIRIH_ERR_NCCL;
  ncclUniqueId uid;
  if (nccl_dyn()->GetUniqueId(&uid) != ncclSuccess) return GIRIH_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) <= GIRIH_COMM_ID_BYTES, "id size");
  memset(id, 0, len);
  memcpy(id, &uid, sizeof(uid));
  return GIRIH_OK;
}

extern "C" int girih_gpu_comm_init(girih_gpu_ctx *c, const void *id, size_t len) {
  if (!c || !id || len < sizeof(ncclUniqueId)) return GIRIH_ERR_ARG;
  if (c->nranks == 1) return GIRIH_OK;
  if (!nccl_dyn()) return fail(c, GIRIH_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", nccl_dyn_error());
  CU(cudaSetDevice(c->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  NC(nccl_dyn()->CommInitRank(&c->comm, c->nranks, uid, c->rank));
  // the thinnest slab bounds every halo depth; all ranks must derive the same exchange schedule from it
  int *d_nz = reinterpret_cast<int *>(c->d_scan);
  CU(cudaMemcpyAsync(d_nz, &c->g.nz, sizeof(int), cudaMemcpyHostToDevice, c->s_comm));
  NC(nccl_dyn()->AllReduce(d_nz, d_nz + 1, 1, ncclInt32, ncclMin, c->comm, c->s_comm));
  CU(cudaMemcpyAsync(&c->nz_min, d_nz + 1, sizeof(int), cudaMemcpyDeviceToHost, c->s_comm));
  CU(cudaStreamSynchronize(c->s_comm));
  c->halo_max = std::min(c->g.Z0, std::max(c->nz_min, c->g.r));
  return GIRIH_OK;
}

extern "C" int girih_plan_halo_exchange(int nz, int depth, int rank, int nranks, int *send_down, int *recv_down,
                                        int *send_up, int *recv_up);

// Exchange `depth` planes of `arr` with both z neighbours on the comm stream: my top `depth`
// interior planes go to the upper neighbour's lower halo and vice versa (geometry of
// src/mpi_utils.c:173-200 generalised from r to depth = T*r planes).  Planes are contiguous,
// so no pack/unpack kernel is needed.
static int exchange_z(girih_gpu_ctx *c, void *arr, int depth, cudaStream_t s) {
  if (c->dims[2] == 1) return GIRIH_OK;
  if (!c->comm) return fail(c, GIRIH_ERR_STATE, "girih_gpu_comm_init was not called");
  const DevGrid &g = c->g;
  const size_t plane_b = (size_t)g.pxy * c->es;
  const size_t count = (size_t)depth * plane_b;
  char *base = (char *)arr;
  const int up = neighbour(c, 2, +1), dn = neighbour(c, 2, -1);
  int sd, rd, su, ru;
  if (girih_plan_halo_exchange(g.nz, depth, c->coords[2], c->dims[2], &sd, &rd, &su, &ru) != GIRIH_OK)
    return fail(c, GIRIH_ERR_ARG, "halo depth %d does not fit a slab of %d planes", depth, g.nz);
  NC(nccl_dyn()->GroupStart());
  if (dn >= 0) {
    NC(nccl_dyn()->Send(base + (size_t)(g.Z0 + sd) * plane_b, count, ncclChar, dn, c->comm, s));
    NC(nccl_dyn()->Recv(base + (size_t)(g.Z0 + rd) * plane_b, count, ncclChar, dn, c->comm, s));
  }
  if (up >= 0) {
    NC(nccl_dyn()->Send(base + (size_t)(g.Z0 + su) * plane_b, count, ncclChar, up, c->comm, s));
    NC(nccl_dyn()->Recv(base + (size_t)(g.Z0 + ru) * plane_b, count, ncclChar, up, c->comm, s));
  }
  NC(nccl_dyn()->GroupEnd());
  return GIRIH_OK;
}

// ---- x / y faces (src/mpi_utils.c:116-170; the reference's sub_array_copy pack/unpack, :31-45) ----------------
// Copies the box [x0, x0+ex) x [y0, y0+ey) x [z0, z0+ez) (device coordinates) of `arr` into / out of a dense buffer.
template <typename R, bool PACK>
__global__ void k_face(DevGrid g, R *__restrict__ arr, R *__restrict__ buf, int x0, int y0, int z0, int ex, int ey, int ez) {
  const long long n = (long long)ex * ey * ez;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % ex);
    const long long t = i / ex;
    const int y = (int)(t % ey), z = (int)(t / ey);
    R *p = arr + ((long long)(z0 + z) * g.ny_dev + (y0 + y)) * g.px + (x0 + x);
    if (PACK) buf[i] = *p;
    else *p = buf[i];
  }
}

template <bool PACK>
static cudaError_t launch_face(girih_gpu_ctx *c, void *arr, void *buf, int x0, int y0, int z0, int ex, int ey, int ez,
                               cudaStream_t s) {
  const long long n = (long long)ex * ey * ez;
  const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  if (c->es == 8) {
    auto kfn = k_face<double, PACK>;
    GIRIH_LAUNCH(kfn, grid, 256, 0, s, c->g, (double *)arr, (double *)buf, x0, y0, z0, ex, ey, ez);
  } else {
    auto kfn = k_face<float, PACK>;
    GIRIH_LAUNCH(kfn, grid, 256, 0, s, c->g, (float *)arr, (float *)buf, x0, y0, z0, ex, ey, ez);
  }
  c->n_kernels++;
  return cudaGetLastError();
}

// r-deep faces of `arr` with the x neighbours, then with the y neighbours.  The y faces span the x halos just
// received and the z planes exchanged afterwards span both, so edge and corner cells arrive too (the box
// operator reads them; the star operators do not).
static int exchange_xy(girih_gpu_ctx *c, void *arr, cudaStream_t s) {
  if (!xy_decomposed(c)) return GIRIH_OK;
  if (!c->comm) return fail(c, GIRIH_ERR_STATE, "girih_gpu_comm_init was not called");
  const DevGrid &g = c->g;
  const int r = g.r;
  const size_t xface = (size_t)r * g.ny * g.nz, yface = (size_t)(g.nx + 2 * r) * r * g.nz;
  const size_t need = std::max(xface, yface);
  if (c->pack_elems < need) {
    if (c->d_pack) CU(cudaFree(c->d_pack));
    c->d_pack = nullptr;
    CU(cudaMalloc(&c->d_pack, 4 * need * c->es));
    c->pack_elems = need;
  }
  char *buf = (char *)c->d_pack;
  const size_t slot = c->pack_elems * c->es;   // [0] send-, [1] send+, [2] recv-, [3] recv+
  for (int d = 0; d < 2; ++d) {
    if (c->dims[d] == 1) continue;
    const int lo = neighbour(c, d, -1), hi = neighbour(c, d, +1);
    // extent of one face and the device coordinates of: my lowest / highest interior layers, the halo layers
    const int ex = d == 0 ? r : g.nx + 2 * r, ey = d == 0 ? g.ny : r, ez = g.nz;
    const int xs = d == 0 ? g.X0 : g.X0 - r, ys = g.Y0;
    const int send_lo[2] = {xs, ys}, send_hi[2] = {d == 0 ? g.X0 + g.nx - r : xs, d == 1 ? g.Y0 + g.ny - r : ys};
    const int recv_lo[2] = {d == 0 ? g.X0 - r : xs, d == 1 ? g.Y0 - r : ys};
    const int recv_hi[2] = {d == 0 ? g.X0 + g.nx : xs, d == 1 ? g.Y0 + g.ny : ys};
    const size_t bytes = (size_t)ex * ey * ez * c->es;
    if (lo >= 0) CU(launch_face<true>(c, arr, buf + 0 * slot, send_lo[0], send_lo[1], g.Z0, ex, ey, ez, s));
    if (hi >= 0) CU(launch_face<true>(c, arr, buf + 1 * slot, send_hi[0], send_hi[1], g.Z0, ex, ey, ez, s));
    NC(nccl_dyn()->GroupStart());
    if (lo >= 0) {
      NC(nccl_dyn()->Send(buf + 0 * slot, bytes, ncclChar, lo, c->comm, s));
      NC(nccl_dyn()->Recv(buf + 2 * slot, bytes, ncclChar, lo, c->comm, s));
    }
    if (hi >= 0) {
      NC(nccl_dyn()->Send(buf + 1 * slot, bytes, ncclChar, hi, c->comm, s));
      NC(nccl_dyn()->Recv(buf + 3 * slot, bytes, ncclChar, hi, c->comm, s));
    }
    NC(nccl_dyn()->GroupEnd());
    if (lo >= 0) CU(launch_face<false>(c, arr, buf + 2 * slot, recv_lo[0], recv_lo[1], g.Z0, ex, ey, ez, s));
    if (hi >= 0) CU(launch_face<false>(c, arr, buf + 3 * slot, recv_hi[0], recv_hi[1], g.Z0, ex, ey, ez, s));
  }
  return GIRIH_OK;
}

static int timed_exchange(girih_gpu_ctx *c, void *arr, int depth) {
  // runs on the comm stream, ordered after everything issued so far on the compute stream, and
  // makes the compute stream wait for its completion
  if (c->nranks == 1) return GIRIH_OK;
  if (c->comm_ev_used == c->comm_ev.size()) {
    EvPair p;
    CU(cudaEventCreate(&p.a));
    CU(cudaEventCreate(&p.b));
    c->comm_ev.push_back(p);
  }
  EvPair &p = c->comm_ev[c->comm_ev_used++];
  CU(cudaEventRecord(c->ev_x, c->s_comp));
  CU(cudaStreamWaitEvent(c->s_comm, c->ev_x, 0));
  CU(cudaEventRecord(p.a, c->s_comm));
  int rc = exchange_xy(c, arr, c->s_comm);
  if (rc) return rc;
  if ((rc = exchange_z(c, arr, depth, c->s_comm))) return rc;
  CU(cudaEventRecord(p.b, c->s_comm));
  CU(cudaStreamWaitEvent(c->s_comp, p.b, 0));
  return GIRIH_OK;
}

// time-invariant arrays (roc2, per-point coefficients) need their deep halos once
static int exchange_static(girih_gpu_ctx *c) {
  if (c->nranks == 1 || c->static_halo_done) return GIRIH_OK;
  int rc = 0;
  if (c->dU3 && (rc = timed_exchange(c, c->dU3, c->halo_max))) return rc;
  for (int m = 0; m < c->kd.n_coef_arrays; ++m)
    if ((rc = timed_exchange(c, (char *)c->dCoef + (size_t)m * c->arr_elems * c->es, c->halo_max))) return rc;
  c->static_halo_done = true;
  return GIRIH_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------------
template <typename R> static ConstCoef<R> make_cc(const girih_gpu_ctx *c) {
  ConstCoef<R> k;
  for (int i = 0; i < 5; ++i) k.v[i] = (R)c->cc[i];
  return k;
}

// ---- naive: one step over a device-coordinate box -------------------------------------------
template <int K, typename R, bool FM>
static cudaError_t launch_naive_t(girih_gpu_ctx *c, int dst, int xb, int yb, int zb, int xe, int ye, int ze) {
  if (xe <= xb || ye <= yb || ze <= zb) return cudaSuccess;
  dim3 block(64, 4, 1);
  dim3 grid((xe - xb + 63) / 64, (ye - yb + 3) / 4, ze - zb);
  R *u = (R *)c->dU[dst];
  const R *v = (const R *)c->dU[dst ^ 1];
  auto kfn = k_naive<K, R, FM>;
  GIRIH_LAUNCH(kfn, grid, block, 0, c->s_comp, c->g, u, v, (const R *)c->dU3, (const R *)c->dCoef,
               (long long)c->arr_elems, make_cc<R>(c), xb, yb, zb, xe, ye, ze);
  c->n_kernels++;
  return cudaGetLastError();
}
static cudaError_t launch_naive(girih_gpu_ctx *c, int dst, int xb, int yb, int zb, int xe, int ye, int ze) {
#define GN(K)                                                                          \
  case K:                                                                              \
    if (c->opt_contract)                                                               \
      return c->es == 8 ? launch_naive_t<K, double, true>(c, dst, xb, yb, zb, xe, ye, ze)  \
                        : launch_naive_t<K, float, true>(c, dst, xb, yb, zb, xe, ye, ze);  \
    return c->es == 8 ? launch_naive_t<K, double, false>(c, dst, xb, yb, zb, xe, ye, ze)   \
                      : launch_naive_t<K, float, false>(c, dst, xb, yb, zb, xe, ye, ze);
  switch (c->kernel) { GN(0) GN(1) GN(2) GN(3) GN(4) GN(5) GN(7) default: return cudaErrorInvalidValue; }
#undef GN
}

// One fused pass: T steps reading array `src` and writing array `dst` (src != dst) on the output
// planes [zb0, ze0) (device z).  T == 1 is the single step of ts 0/1.
// A second range [zb1, ze1) is swept by the same launch where the kernel supports it (fused r = 1 sweep),
// else by a second launch.
static cudaError_t launch_pass(girih_gpu_ctx *c, int T, int src, int dst, int zb0, int ze0, int zb1 = 0, int ze1 = 0) {
  const DevGrid &g = c->g;
  if (ze0 <= zb0) return cudaSuccess;
  if (ze1 > zb1) {
    const bool fused_kernel = (g.r == 1) && (c->kernel != 7) && (c->opt_variant != 1) && (T > 1 || c->opt_variant == 2);
    if (!fused_kernel) {
      cudaError_t e = launch_pass(c, T, src, dst, zb0, ze0);
      return e != cudaSuccess ? e : launch_pass(c, T, src, dst, zb1, ze1);
    }
  }
  // variant 0 picks the fastest measured single-step kernel per operator and precision
  // (profiles/kernel_sweep_r01.md): the marching kernel, except for the fp64 variable-coefficient
  // operators where one thread per site with all loads in flight runs at the HBM limit already.
  bool naive = (c->opt_variant == 1);
  if (c->opt_variant == 0 && T == 1 && c->es == 8 && (c->kernel == 2 || c->kernel == 3 || c->kernel == 5)) naive = true;
  if (naive) {
    if (T != 1) return cudaErrorInvalidValue;
    return launch_naive(c, dst, g.X0, g.Y0, zb0, g.X0 + g.nx, g.Y0 + g.ny, ze0);
  }
  StreamLaunch sl;
  sl.g = g;
  sl.in = c->dU[src];
  sl.out = c->dU[dst];
  sl.roc2 = c->dU3;
  sl.coef = c->dCoef;
  sl.coef_stride = (long long)c->arr_elems;
  for (int i = 0; i < 5; ++i) sl.cc[i] = c->cc[i];
  sl.zb0 = zb0;
  sl.ze0 = ze0;
  sl.zb1 = zb1;
  sl.ze1 = ze1;
  sl.zchunk = c->opt_zchunk;
  sl.tile = c->opt_tile ? c->opt_tile : c->tuned_tile[T & 7];
  sl.variant = c->opt_variant;
  sl.contract = c->opt_contract;
  sl.stream = c->s_comp;
  if (tile_is_exact(sl.tile) && g.r == 1 && T > 1) {
    if (!c->d_xbuf) {   // 48 KB of edge slots per resident CTA, zeroed once: tag 0 is never used
      cudaDeviceGetAttribute(&c->nsm, cudaDevAttrMultiProcessorCount, c->device);
      const size_t bytes = (size_t)c->nsm * 48 * 1024;
      if (cudaMalloc((void **)&c->d_xbuf, bytes) == cudaSuccess && cudaMalloc((void **)&c->d_xerr, sizeof(int)) == cudaSuccess) {
        cudaMemsetAsync(c->d_xbuf, 0, bytes, c->s_comp);
        cudaMemsetAsync(c->d_xerr, 0, sizeof(int), c->s_comp);
        c->xbuf_bytes = bytes;
        c->xseq = 0;
      } else {
        (void)cudaGetLastError();
      }
    }
    if (c->xseq > 0xf0000000u) {   // tags are 32 bits: start over from clean slots long before they wrap
      cudaMemsetAsync(c->d_xbuf, 0, c->xbuf_bytes, c->s_comp);
      c->xseq = 0;
    }
    sl.xbuf = c->xbuf_bytes ? c->d_xbuf : nullptr;
    sl.xbuf_bytes = c->xbuf_bytes;
    sl.xseq = &c->xseq;
    sl.xerr = c->d_xerr;
    sl.nsm = c->nsm;
    c->xbuf_used = c->xbuf_bytes != 0;
  }
  if (c->push_planes > 0) {   // this pass also stores its boundary planes into the neighbours' halos (run_passes)
    const size_t plane_b = (size_t)g.pxy * c->es;
    if (c->peer_U[1][dst]) {   // my planes [Z0+nz-D, Z0+nz) -> upper neighbour's planes [Z0-D, Z0): shift by -nz
      sl.push_up = (char *)c->peer_U[1][dst] - (size_t)g.nz * plane_b;
      sl.push_up_from = g.Z0 + g.nz - c->push_planes;
    }
    if (c->peer_U[0][dst]) {   // my planes [Z0, Z0+D) -> lower neighbour's planes [Z0+nz', Z0+nz'+D): shift by +nz'
      sl.push_dn = (char *)c->peer_U[0][dst] + (size_t)c->peer_nz[0] * plane_b;
      sl.push_dn_below = g.Z0 + c->push_planes;
    }
  }
  c->n_kernels++;
  if (c->kernel == 7) return T == 1 ? launch_box(c->es, sl) : cudaErrorInvalidValue;
  if (g.r == 1) {
    const unsigned seq_before = c->xseq;
    const cudaError_t e = launch_r1(c->kernel, c->es, T, sl);
    if (T > 1) c->n_fused++;
    if (c->xseq != seq_before) c->n_exact++;   // the exact-tile launcher consumed slot tags: it ran
    return e;
  }
  if (T != 1) return cudaErrorInvalidValue;
  return launch_r4(c->kernel, c->es, sl);
}

// ------------------------------------------------------------------------------------------------
// steppers
// ------------------------------------------------------------------------------------------------
// Fusion depth used when the caller does not ask for one: the fastest measured depth per operator
// and precision on B200 at 512^3 (profiles/kernel_sweep_r01.md).  Constant coefficients gain most;
// with per-point coefficient arrays the coefficient planes dominate the traffic and are re-read by
// every fused level, so the gain shrinks (slot 2, 3) or vanishes (slot 5: 9 words per update).
static int default_tfuse(const girih_gpu_ctx *c) {
  switch (c->kernel) {
    // fastest measured depth per (operator, precision) at 512^3: profiles/kernel_sweep_r01.md and the round-2 sweep
    // profiles/r02_kbench_others.log (with the trapezoid-skip tile the fp64 per-point-coefficient operators gain from
    // depth 2: slot 3 168.8 vs 145.6 GLUP/s, slot 5 105.8 vs 98.8, slot 2 256.4 vs 192.5; slot 3 fp32 353.8 at depth 3)
    case 1: return 4;
    case 2: return c->es == 8 ? 2 : 3;
    case 3: return c->es == 8 ? 2 : 3;
    case 5: return 2;
    default: return 1;
  }
}

static int begin_run(girih_gpu_ctx *c) {
  if (!c->uploaded) return fail(c, GIRIH_ERR_STATE, "run before upload");
  CU(cudaSetDevice(c->device));
  c->n_kernels = c->n_passes = c->n_steps = 0;
  c->comm_ev_used = 0;
  c->comp_ev_used = 0;
  c->ms_compute = c->ms_comm = c->ms_total = 0;
  if (c->d_flags) CU(cudaMemsetAsync(c->d_flags + 2, 0, sizeof(int), c->s_comp));   // "a halo-push wait gave up" is per run
  CU(cudaEventRecord(c->ev_t0, c->s_comp));
  return GIRIH_OK;
}
static int end_run(girih_gpu_ctx *c) {
  CU(cudaEventRecord(c->ev_t1, c->s_comp));
  CU(cudaStreamSynchronize(c->s_comp));
  CU(cudaStreamSynchronize(c->s_comm));
  if (c->xbuf_used) {   // did a tile of the exact-tiled sweep give up on a neighbour tile?
    int gave_up = 0;
    CU(cudaMemcpy(&gave_up, c->d_xerr, sizeof(int), cudaMemcpyDeviceToHost));
    if (gave_up) {
      CU(cudaMemset(c->d_xerr, 0, sizeof(int)));
      return fail(c, GIRIH_ERR_STATE, "exact-tiled sweep: a tile waited for a neighbour's edge values beyond the limit");
    }
  }
  if (c->d_flags && c->opt_push) {   // did a halo-push wait give up on a neighbour?
    int gave_up = 0;
    CU(cudaMemcpy(&gave_up, c->d_flags + 2, sizeof(int), cudaMemcpyDeviceToHost));
    if (gave_up) return fail(c, GIRIH_ERR_STATE, "halo push: a neighbour did not signal within the time limit");
  }
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
  c->ms_total = ms;
  double comm = 0;
  for (size_t i = 0; i < c->comm_ev_used; ++i) {
    float m = 0;
    CU(cudaEventElapsedTime(&m, c->comm_ev[i].a, c->comm_ev[i].b));
    comm += m;
  }
  c->ms_comm = comm;
  double comp = 0;   // time the compute stream spent inside sweeps (waits for exchanges / neighbours are outside the pairs)
  for (size_t i = 0; i < c->comp_ev_used; ++i) {
    float m = 0;
    CU(cudaEventElapsedTime(&m, c->comp_ev[i].a, c->comp_ev[i].b));
    comp += m;
  }
  c->ms_compute = c->comp_ev_used ? comp : c->ms_total - comm;
  return GIRIH_OK;
}

// Runs the passes in `sizes` (steps per pass) back to back over all local interior planes.  Every
// pass reads the array holding the newest level (`cur`) and writes the other one.
// overlap: the planes the neighbours need for the NEXT pass are computed first and their exchange
// runs on the comm stream underneath the rest of the pass (the halo-first pattern,
// src/kernels/halo_first_ts.c:156-194, with depth T*r instead of r).
static void plan_passes(int nsteps, int T, std::vector<int> &sizes);

// Exchange schedule of a run: depth[p] = halo planes exchanged before pass p (0 = none).  One exchange can
// serve up to `group` consecutive passes: it brings depth = sum of their T*r planes, and every pass but the
// last of the group also computes the planes of the neighbour's slab that the following passes will read
// (the same deep-halo recomputation a fused pass does for its own T levels), so fewer, larger messages and
// fewer points where the slabs wait for each other.
static void plan_exchanges(const std::vector<int> &sizes, int r, int cap, int group, std::vector<int> &depth) {
  depth.assign(sizes.size(), 0);
  int ready = 0;
  for (size_t p = 0; p < sizes.size(); ++p) {
    const int need = sizes[p] * r;
    if (ready < need) {
      int D = need, cnt = 1;
      for (size_t q = p + 1; q < sizes.size() && cnt < group && D + sizes[q] * r <= cap; ++q, ++cnt) D += sizes[q] * r;
      depth[p] = D;
      ready = D;
    }
    ready -= need;
  }
}

extern "C" int girih_plan_fused_exchanges(int nsteps, int tfuse, int r, int halo_cap, int group, int *depth,
                                          int max_n, int *n) {
  if (nsteps < 0 || tfuse < 1 || r < 1 || halo_cap < tfuse * r || group < 1 || !n) return GIRIH_ERR_ARG;
  std::vector<int> sizes, d;
  plan_passes(nsteps, tfuse, sizes);
  plan_exchanges(sizes, r, halo_cap, group, d);
  *n = (int)d.size();
  if (depth) {
    if ((int)d.size() > max_n) return GIRIH_ERR_ARG;
    for (size_t i = 0; i < d.size(); ++i) depth[i] = d[i];
  }
  return GIRIH_OK;
}

// passes served by one exchange: the option, else up to 4 while the recomputed planes stay a small
// fraction of the thinnest slab
static int halo_group(const girih_gpu_ctx *c, int T, bool overlap) {
  if (c->nranks == 1 || overlap || c->kd.time_order != 1 || xy_decomposed(c)) return 1;
  if (c->opt_halo_group > 0) return c->opt_halo_group;
  int k = 1;
  while (k < 4 && (k + 1) * T * c->g.r <= c->halo_max && k * T * c->g.r * 16 <= c->nz_min) ++k;
  return k;
}

// launch_pass bracketed by a cudaEvent pair on the compute stream (the run's "compute" time is the sum of the pairs)
static cudaError_t timed_pass(girih_gpu_ctx *c, int T, int src, int dst, int zb0, int ze0, int zb1 = 0, int ze1 = 0) {
  if (c->comp_ev_used == c->comp_ev.size()) {
    EvPair p;
    cudaError_t e = cudaEventCreate(&p.a);
    if (e == cudaSuccess) e = cudaEventCreate(&p.b);
    if (e != cudaSuccess) return e;
    c->comp_ev.push_back(p);
  }
  EvPair &p = c->comp_ev[c->comp_ev_used++];
  cudaError_t e = cudaEventRecord(p.a, c->s_comp);
  if (e == cudaSuccess) e = launch_pass(c, T, src, dst, zb0, ze0, zb1, ze1);
  if (e == cudaSuccess) e = cudaEventRecord(p.b, c->s_comp);
  return e;
}

static int run_passes(girih_gpu_ctx *c, const std::vector<int> &sizes, int &cur, bool overlap) {
  const DevGrid &g = c->g;
  const int r = g.r;
  const int zb = g.Z0, ze = g.Z0 + g.nz;
  int ready = 0;   // depth of valid halo planes already present around array `cur`
  int rc;
  int Tmax = 1;
  for (int s : sizes) Tmax = std::max(Tmax, s);
  if (xy_decomposed(c)) overlap = false;   // x/y faces are exchanged between whole steps only
  const bool copy = copy_enabled(c);       // halo copy: the overlapped schedule, halos moved by the copy engines
  if (copy) overlap = true;
  const bool push = !overlap && push_enabled(c, Tmax);   // halo push: the sweep itself feeds the neighbours' halos
  const int group = push ? 1 : halo_group(c, Tmax, overlap);
  std::vector<int> depth;
  plan_exchanges(sizes, r, c->halo_max, group, depth);
  bool copy_live = false;   // the halos of `cur` were delivered by the neighbours' copies (wait for their flags)
  if (group > 1) {
    // the Dirichlet frame cells of the deep-halo planes are never written by a kernel: bring them (with
    // the planes) into BOTH arrays once; later exchanges and extended sweeps keep them
    if ((rc = timed_exchange(c, c->dU[0], c->halo_max))) return rc;
    if ((rc = timed_exchange(c, c->dU[1], c->halo_max))) return rc;
  }
  for (size_t p = 0; p < sizes.size(); ++p) {
    const int T = sizes[p];
    const int src = cur, dst = cur ^ 1;
    if (push) {
      const bool need_dn = neighbour(c, 2, -1) >= 0, need_up = neighbour(c, 2, +1) >= 0;
      auto signal = [&]() -> cudaError_t {   // "everything I issued so far on the compute stream is done"
        c->push_seq++;
        c->n_kernels++;
        GIRIH_LAUNCH(k_flag_signal, 1, 1, 0, c->s_comp, (volatile int *)(need_dn ? c->peer_flags[0] + 1 : nullptr),
                     (volatile int *)(need_up ? c->peer_flags[1] + 0 : nullptr), c->push_seq);
        return cudaGetLastError();
      };
      if (p == 0) {
        // Level-0 halos come through NCCL once per run -- both arrays, because the Dirichlet frame cells of the halo
        // planes are never pushed (a sweep stores interior points only).  The flag behind the exchange tells the
        // neighbours that my receives have landed, so none of their pushes can be overtaken by a late receive.
        if ((rc = timed_exchange(c, c->dU[dst], Tmax * r))) return rc;
        if ((rc = timed_exchange(c, c->dU[src], T * r))) return rc;
        CU(signal());
      }
      // both neighbours have passed their last signal (end of pass p-1, or of the initial exchange): their stores
      // into my halos are complete and they no longer read the halo planes this pass overwrites
      GIRIH_LAUNCH(k_flag_wait, 1, 1, 0, c->s_comp, (volatile int *)c->d_flags, (int)need_dn, (int)need_up, c->push_seq,
                   flag_wait_cycles());
      CU(cudaGetLastError());
      c->n_kernels++;
      c->push_planes = (p + 1 < sizes.size()) ? sizes[p + 1] * r : 0;   // what the next pass reads beyond its slab
      cudaError_t le = timed_pass(c, T, src, dst, zb, ze);
      c->push_planes = 0;
      CU(le);
      CU(signal());
      ready = 0;
      cur = dst;
      c->n_passes++;
      c->n_steps += T;
      continue;
    }
    if (c->nranks > 1 && !overlap) {
      if (depth[p] > 0) {
        if ((rc = timed_exchange(c, c->dU[src], depth[p]))) return rc;
        ready = depth[p];
      }
      // planes beyond the slab that later passes of this group read: computed here, towards neighbours only
      const int ext = ready - T * r;
      const int lo = (neighbour(c, 2, -1) >= 0) ? ext : 0, hi = (neighbour(c, 2, +1) >= 0) ? ext : 0;
      CU(timed_pass(c, T, src, dst, zb - lo, ze + hi));
      ready = ext;
      cur = dst;
      c->n_passes++;
      c->n_steps += T;
      continue;
    }
    if (c->nranks > 1 && ready < T * r) {
      if ((rc = timed_exchange(c, c->dU[src], T * r))) return rc;
    }
    if (copy_live) {
      // the halos of `src` were copied in by the neighbours during their previous pass: wait for both flags
      const bool wdn = neighbour(c, 2, -1) >= 0, wup = neighbour(c, 2, +1) >= 0;
      GIRIH_LAUNCH(k_flag_wait, 1, 1, 0, c->s_comp, (volatile int *)c->d_flags, (int)wdn, (int)wup, c->push_seq, flag_wait_cycles());
      CU(cudaGetLastError());
      c->n_kernels++;
      copy_live = false;
    }
    const int nd = (p + 1 < sizes.size()) ? sizes[p + 1] * r : r;   // halo depth the next pass needs
    // (decided on the thinnest slab of the run so that every rank takes the same branch: the two branches
    // exchange at different points)
    // halo copy: the outer parts must be at least T*r planes thick, so that the sweep of the inner part never reads a halo
    // plane -- the neighbours' NEXT copies land in the halo planes of this pass's source array as soon as they have my
    // flag, i.e. possibly while my inner sweep is still running (found by the emulator fuzzer on 4- and 9-plane slabs)
    const int nq = copy ? std::max(nd, T * r) : nd;
    if (overlap && c->nranks > 1 && c->nz_min >= 4 * nq) {
      // the two outer quarters of the slab first (one launch, full waves: thin boundary launches would pay
      // the 2T-plane pipeline fill for a few planes), then the halo exchange of the new level runs under
      // the sweep of the inner half
      const int zq = std::max(nq, g.nz / 4);
      const bool need_dn = neighbour(c, 2, -1) >= 0, need_up = neighbour(c, 2, +1) >= 0;
      CU(timed_pass(c, T, src, dst, zb, zb + zq, ze - zq, ze));
      CU(cudaEventRecord(c->ev_y, c->s_comp));
      CU(cudaStreamWaitEvent(c->s_comm, c->ev_y, 0));
      if (c->comm_ev_used == c->comm_ev.size()) {
        EvPair q;
        CU(cudaEventCreate(&q.a));
        CU(cudaEventCreate(&q.b));
        c->comm_ev.push_back(q);
      }
      EvPair &q = c->comm_ev[c->comm_ev_used++];
      CU(cudaEventRecord(q.a, c->s_comm));
      if (copy) {
        const size_t plane_b = (size_t)g.pxy * c->es, bytes = (size_t)nd * plane_b;
        char *mine = (char *)c->dU[dst];
        if (need_up)   // my top nd planes -> the upper neighbour's planes [Z0 - nd, Z0)
          CU(cudaMemcpyAsync((char *)c->peer_U[1][dst] + (size_t)(g.Z0 - nd) * plane_b, mine + (size_t)(ze - nd) * plane_b, bytes,
                             cudaMemcpyDeviceToDevice, c->s_comm));
        if (need_dn)   // my lowest nd planes -> the lower neighbour's planes [Z0 + nz', Z0 + nz' + nd)
          CU(cudaMemcpyAsync((char *)c->peer_U[0][dst] + (size_t)(g.Z0 + c->peer_nz[0]) * plane_b, mine + (size_t)zb * plane_b, bytes,
                             cudaMemcpyDeviceToDevice, c->s_comm));
        c->push_seq++;
        c->n_kernels++;
        GIRIH_LAUNCH(k_flag_signal, 1, 1, 0, c->s_comm, (volatile int *)(need_dn ? c->peer_flags[0] + 1 : nullptr),
                     (volatile int *)(need_up ? c->peer_flags[1] + 0 : nullptr), c->push_seq);
        CU(cudaGetLastError());
        copy_live = true;
      } else {
        if ((rc = exchange_z(c, c->dU[dst], nd, c->s_comm))) return rc;
      }
      CU(cudaEventRecord(q.b, c->s_comm));
      CU(timed_pass(c, T, src, dst, zb + zq, ze - zq));
      CU(cudaStreamWaitEvent(c->s_comp, q.b, 0));   // my next-but-one outer sweep overwrites what this exchange reads
      ready = nd;
    } else {
      CU(timed_pass(c, T, src, dst, zb, ze));
      ready = 0;
    }
    cur = dst;
    c->n_passes++;
    c->n_steps += T;
  }
  return GIRIH_OK;
}

// The pass schedule of the fused stepper.  All but the last step are fused; the last one is a single
// step so that BOTH arrays end up holding the levels the reference leaves (newest and newest-1).
// Every pass moves the newest level to the other array, and level n must end in U1 when n is odd: the
// number of passes that cover the first nsteps-1 steps must have the parity of nsteps-1.
static void plan_passes(int nsteps, int T, std::vector<int> &sizes) {
  sizes.clear();
  if (nsteps <= 0) return;
  if (T < 1) T = 1;
  int n1 = nsteps - 1;
  while (n1 >= T) { sizes.push_back(T); n1 -= T; }
  if (n1 > 0) sizes.push_back(n1);
  if (((int)sizes.size() - (nsteps - 1)) % 2 != 0) {
    // some pass has >= 2 steps here (otherwise the count equals the step count): split it
    for (size_t i = sizes.size(); i-- > 0;)
      if (sizes[i] >= 2) {
        const int a1 = sizes[i] / 2, a2 = sizes[i] - a1;
        sizes[i] = a2;
        sizes.insert(sizes.begin() + (long)i + 1, a1);
        break;
      }
  }
  sizes.push_back(1);
}

extern "C" int girih_plan_fused_passes(int nsteps, int tfuse, int *sizes, int max_sizes, int *n_sizes) {
  if (nsteps < 0 || tfuse < 1 || !n_sizes) return GIRIH_ERR_ARG;
  std::vector<int> v;
  plan_passes(nsteps, tfuse, v);
  *n_sizes = (int)v.size();
  if (sizes) {
    if ((int)v.size() > max_sizes) return GIRIH_ERR_ARG;
    for (size_t i = 0; i < v.size(); ++i) sizes[i] = v[i];
  }
  return GIRIH_OK;
}

// Plane ranges of one z-halo exchange of `depth` planes, in LOCAL plane coordinates (0 = first interior
// plane of the slab, nz = number of interior planes).  A negative start means "no neighbour on that side".
extern "C" int girih_plan_halo_exchange(int nz, int depth, int rank, int nranks, int *send_down, int *recv_down,
                                        int *send_up, int *recv_up) {
  if (nz < 1 || depth < 1 || depth > nz || nranks < 1 || rank < 0 || rank >= nranks) return GIRIH_ERR_ARG;
  const bool dn = rank > 0, up = rank + 1 < nranks;
  if (send_down) *send_down = dn ? 0 : -1;               // my lowest `depth` interior planes
  if (recv_down) *recv_down = dn ? -depth : -1 - depth;   // land below my interior
  if (send_up) *send_up = up ? nz - depth : -1;          // my highest `depth` interior planes
  if (recv_up) *recv_up = up ? nz : -1;                  // land above my interior
  return GIRIH_OK;
}

static int finish_halos(girih_gpu_ctx *c) {
  // leave both arrays with exchanged r-deep halos, as the reference's steppers do
  // (src/kernels/nb_naive_ts.c:192,198)
  if (c->nranks == 1) return GIRIH_OK;
  int rc;
  if ((rc = timed_exchange(c, c->dU[0], c->g.r))) return rc;
  if ((rc = timed_exchange(c, c->dU[1], c->g.r))) return rc;
  return GIRIH_OK;
}

// ------------------------------------------------------------------------------------------------
// slot 6 (solar): a time step is the H update of every interior cell followed by the E update
// (src/kernels/solar_spt_blk.ic:388-397), two launches on the compute stream
// ------------------------------------------------------------------------------------------------
static cudaError_t solar_step(girih_gpu_ctx *c, int xb, int yb, int zb, int xe, int ye, int ze, int phases) {
  SolarLaunch s;
  s.u = c->dU[0];
  s.coef = c->dCoef;
  s.n2 = c->solar_n2;
  s.nnx = c->hshape[0]; s.nny = c->hshape[1];
  s.xb = xb; s.xe = xe; s.yb = yb; s.ye = ye; s.zb = zb; s.ze = ze;
  s.zchunk = c->opt_zchunk;
  s.tile = c->opt_tile ? c->opt_tile : c->tuned_tile[1];   // option, else what girih_gpu_autotune picked, else the default
  s.stream = c->s_comp;
  for (int ph = 0; ph < 2; ++ph) {
    if (!((phases >> ph) & 1)) continue;
    cudaError_t e = launch_solar(c->es, ph, s);
    if (e != cudaSuccess) return e;
    c->n_kernels++;
  }
  return cudaSuccess;
}
static int solar_check(girih_gpu_ctx *c) {
  if (c->opt_contract) return fail(c, GIRIH_ERR_UNSUPPORTED, "option contract is not offered for the solar slot");
  return GIRIH_OK;
}
static int solar_run(girih_gpu_ctx *c, int nsteps) {
  int rc;
  if ((rc = solar_check(c))) return rc;
  if ((rc = begin_run(c))) return rc;
  for (int s = 0; s < nsteps; ++s) {
    CU(solar_step(c, 1, 1, 1, c->hshape[0] - 1, c->hshape[1] - 1, c->hshape[2] - 1, 3));
    c->n_passes++;
    c->n_steps++;
  }
  c->tfuse_used = 1;
  return end_run(c);
}

extern "C" int girih_gpu_run_single(girih_gpu_ctx *c, int nsteps, int overlap) {
  if (!c || nsteps < 0) return GIRIH_ERR_ARG;
  int rc;
  if (c->solar) return solar_run(c, nsteps);
  if ((rc = begin_run(c))) return rc;
  if ((rc = exchange_static(c))) return rc;
  std::vector<int> sizes((size_t)nsteps, 1);
  int cur = 1;   // level 0 is read from U2 by step 1 (src/kernels/nb_naive_ts.c:189)
  if ((rc = run_passes(c, sizes, cur, overlap != 0))) return rc;
  if ((rc = finish_halos(c))) return rc;
  c->tfuse_used = 1;
  return end_run(c);
}

// z-wavefront temporal blocking THROUGH L2 for the operators that have no fused-sweep kernel (radius 4: slots 0 and 4;
// the box operator): the GPU form of the reference's wavefront (src/kernels/stencils_1wf.ic:37-77: W time steps in
// flight along z, each lagging the previous one by r planes, `kt -= NHALO`) with the 126 MB L2 in the role of the CPU's
// last-level cache.  The slab is swept in blocks of B planes; step s of a group of W steps follows step s-1 at a
// distance of r planes, so what step s reads -- the planes step s-1 has just written and the planes it read -- is still
// in L2: W steps cost about one read and one write of each array from HBM instead of W.  Every launch is the ordinary
// single-step kernel on a z range, so the result is bit-identical by construction.
//   step s of a group reads array cur ^ (s & 1) and writes the other one
//   RAW: step s stops r planes below the last plane step s-1 has completed (or runs to the end once step s-1 is done)
//   WAR: the planes step s overwrites lie below everything step s-1 still has to read (same distance r)
static int run_zwave(girih_gpu_ctx *c, int nsteps, int W, int B, int &cur) {
  const DevGrid &g = c->g;
  const int r = g.r, zb = g.Z0, ze = g.Z0 + g.nz;
  int done = 0;
  std::vector<int> hi;
  while (done < nsteps) {
    const int w = std::min(W, nsteps - done);
    hi.assign((size_t)w, zb);
    while (hi[w - 1] < ze) {
      for (int s = 0; s < w; ++s) {
        const int limit = (s == 0) ? std::min(ze, hi[0] + B) : (hi[s - 1] >= ze ? ze : hi[s - 1] - r);
        if (limit > hi[s]) {
          const int src = cur ^ (s & 1);
          CU(launch_pass(c, 1, src, src ^ 1, hi[s], limit));   // (no event pair per launch: there are thousands)
          hi[s] = limit;
        }
      }
    }
    done += w;
    c->n_passes++;
    c->n_steps += w;
    if (w & 1) cur ^= 1;
  }
  return GIRIH_OK;
}

// steps in flight of the z wavefront for this context (1 = plain single steps)
static int zwave_depth(const girih_gpu_ctx *c) {
  if (c->nranks != 1 || xy_decomposed(c) || c->kd.max_tfuse > 1 || c->opt_variant == 1) return 1;
  if (c->opt_zwave > 0) return c->opt_zwave;
  return 1;
}

extern "C" int girih_gpu_run_fused(girih_gpu_ctx *c, int nsteps, int tfuse) {
  if (!c || nsteps < 0) return GIRIH_ERR_ARG;
  // the reference's default wavefront (mwd_func_list, src/kernels/stencils.c:303-311) is not_supported_mwd for this slot
  if (c->solar) return fail(c, GIRIH_ERR_UNSUPPORTED, "%s", girih_gpu_strerror(GIRIH_ERR_UNSUPPORTED));
  int T = tfuse;
  if (T <= 0) T = c->tuned_tfuse > 0 ? c->tuned_tfuse : default_tfuse(c);
  T = std::min(T, c->kd.max_tfuse);
  if (c->opt_variant == 1 || c->kernel == 7 || xy_decomposed(c)) T = 1;
  CU(cudaSetDevice(c->device));
  if (c->nranks > 1) T = std::min(T, std::max(1, c->nz_min / std::max(1, c->g.r)));   // same depth on every rank
  if (T > 1) { int frc = refresh_frames_equal(c); if (frc) return frc; }
  if (T > 1 && !c->frames_equal) return fail(c, GIRIH_ERR_FRAME, "%s", girih_gpu_strerror(GIRIH_ERR_FRAME));
  int rc;
  if ((rc = begin_run(c))) return rc;
  if ((rc = exchange_static(c))) return rc;
  std::vector<int> sizes;
  plan_passes(nsteps, T, sizes);
  int cur = 1;
  const int W = zwave_depth(c);
  if (T == 1 && W > 1) {
    const int B = c->opt_zwave_block > 0 ? c->opt_zwave_block : 8;
    if ((rc = run_zwave(c, nsteps, W, B, cur))) return rc;
    T = W;
  } else if ((rc = run_passes(c, sizes, cur, c->opt_overlap != 0))) return rc;
  if (nsteps > 0 && cur != ((nsteps % 2 == 1) ? 0 : 1))
    return fail(c, GIRIH_ERR_STATE, "internal: pass schedule left the newest level in the wrong array");
  if ((rc = finish_halos(c))) return rc;
  c->tfuse_used = T;
  return end_run(c);
}

extern "C" int girih_gpu_step_box(girih_gpu_ctx *c, int dst, int xb, int yb, int zb, int xe, int ye, int ze) {
  if (!c || (dst != 1 && dst != 2)) return GIRIH_ERR_ARG;
  if (!c->uploaded) return fail(c, GIRIH_ERR_STATE, "step_box before upload");
  if (c->solar) {   // dst is ignored: the one array is updated in place (solar(), solar_spt_blk.ic:388-397, ALL_FIELDS)
    if (xb < 1 || yb < 1 || zb < 1 || xe > c->hshape[0] - 1 || ye > c->hshape[1] - 1 || ze > c->hshape[2] - 1)
      return fail(c, GIRIH_ERR_ARG, "box must lie inside the interior");
    int rc = solar_check(c);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    CU(solar_step(c, xb, yb, zb, xe, ye, ze, 3));
    CU(cudaStreamSynchronize(c->s_comp));
    return GIRIH_OK;
  }
  const DevGrid &g = c->g;
  const int r = g.r;
  if (xb < r || yb < r || zb < r || xe > g.nx + r || ye > g.ny + r || ze > g.nz + r)
    return fail(c, GIRIH_ERR_ARG, "box must lie inside the interior");
  CU(cudaSetDevice(c->device));
  CU(launch_naive(c, dst - 1, xb - r + g.X0, yb - r + g.Y0, zb - r + g.Z0, xe - r + g.X0, ye - r + g.Y0,
                  ze - r + g.Z0));
  CU(cudaStreamSynchronize(c->s_comp));
  return GIRIH_OK;
}

// `reps` passes of depth tfuse over this slab's interior planes, no halo exchange
static int time_pass_local(girih_gpu_ctx *c, int tfuse, int reps, double *ms_per_pass) {
  if (c->solar) {   // one pass = one time step (H + E)
    int rc = solar_check(c);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    c->n_kernels = 0;
    const int xe = c->hshape[0] - 1, ye = c->hshape[1] - 1, ze = c->hshape[2] - 1;
    CU(solar_step(c, 1, 1, 1, xe, ye, ze, 3));   // warm
    CU(cudaEventRecord(c->ev_t0, c->s_comp));
    for (int i = 0; i < reps; ++i) CU(solar_step(c, 1, 1, 1, xe, ye, ze, 3));
    CU(cudaEventRecord(c->ev_t1, c->s_comp));
    CU(cudaStreamSynchronize(c->s_comp));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
    *ms_per_pass = (double)ms / reps;
    c->tfuse_used = 1;
    return GIRIH_OK;
  }
  int T = std::max(1, std::min(tfuse, c->kd.max_tfuse));
  if (c->opt_variant == 1 || c->kernel == 7) T = 1;
  CU(cudaSetDevice(c->device));
  if (T > 1) { int frc = refresh_frames_equal(c); if (frc) return frc; }
  if (T > 1 && !c->frames_equal) return fail(c, GIRIH_ERR_FRAME, "%s", girih_gpu_strerror(GIRIH_ERR_FRAME));
  c->n_kernels = 0;
  const DevGrid &g = c->g;
  int cur = 1;
  CU(launch_pass(c, T, cur, cur ^ 1, g.Z0, g.Z0 + g.nz));   // warm
  cur ^= 1;
  CU(cudaEventRecord(c->ev_t0, c->s_comp));
  for (int i = 0; i < reps; ++i) {
    CU(launch_pass(c, T, cur, cur ^ 1, g.Z0, g.Z0 + g.nz));
    cur ^= 1;
  }
  CU(cudaEventRecord(c->ev_t1, c->s_comp));
  CU(cudaStreamSynchronize(c->s_comp));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
  *ms_per_pass = (double)ms / reps;
  c->tfuse_used = T;
  return GIRIH_OK;
}

extern "C" int girih_gpu_time_pass(girih_gpu_ctx *c, int tfuse, int reps, double *ms_per_pass) {
  if (!c || reps < 1 || !ms_per_pass) return GIRIH_ERR_ARG;
  if (!c->uploaded) return fail(c, GIRIH_ERR_STATE, "time_pass before upload");
  if (c->nranks != 1) return fail(c, GIRIH_ERR_ARG, "time_pass works on single-slab contexts");
  return time_pass_local(c, tfuse, reps, ms_per_pass);
}

// ------------------------------------------------------------------------------------------------
// on-device tuner: the GPU analogue of auto_tune_params (src/kernels/diamond_utils.c:691-847) --
// measure every (fusion depth, tile) this operator has kernels for on the resident slab, keep the
// fastest, report under the reference's "[AUTO TUNE]" prefix.  The fields evolve while measuring:
// the caller uploads them again afterwards.
// ------------------------------------------------------------------------------------------------
static std::vector<int> tile_candidates(const girih_gpu_ctx *c, int T) {
  if (c->solar) return {1, 2, 3, 4, 5};
  if (c->opt_variant == 1) return {0};
  if (c->kernel == 7) return {0, 4, 8, 108, 116, 208, 216};
  if (c->kernel == 0) return {0, 8, 16, 108};
  if (c->kernel == 4) return {0, 8, 16};
  if (T == 1) {
    if (c->opt_variant == 0 && c->es == 8 && (c->kernel == 2 || c->kernel == 3 || c->kernel == 5)) return {0};
    if (c->opt_variant == 2) return {216, 408};
    return {0, 108, 208, 404, 408};
  }
  // 5xxx = split-barrier variants (profiles/kernel_sweep_r01.md, round 1b: +3..8% at T = 2, 3 and in fp32)
  if (c->opt_contract) return c->kernel == 1 ? std::vector<int>{0, 5408, 5216, 9408, 9216} : std::vector<int>{0};
  if (c->kernel == 1) return {216, 408, 312, 310, 316, 5408, 5216, 9408, 9216};
  return {216, 408, 9216, 9408};
}

extern "C" int girih_gpu_autotune(girih_gpu_ctx *c, int fused, int verbose, int *best_tfuse, int *best_tile,
                                  double *best_mlups) {
  if (!c) return GIRIH_ERR_ARG;
  if (!c->uploaded) return fail(c, GIRIH_ERR_STATE, "autotune before upload");
  const int saved_tile = c->opt_tile;
  const double lups = (double)c->g.nx * c->g.ny * c->g.nz;
  int Tmax = fused ? c->kd.max_tfuse : 1;
  { int frc = refresh_frames_equal(c); if (frc) return frc; }
  if (c->opt_variant == 1 || c->kernel == 7 || !c->frames_equal) Tmax = 1;
  if (c->nranks > 1) Tmax = std::min(Tmax, std::max(1, c->nz_min / std::max(1, c->g.r)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, c->s_comp));
  double best = -1;
  int bT = 1, rc = GIRIH_OK;
  if (verbose) printf("[AUTO TUNE] GPU kernels of operator %d (%s): fusion depth 1..%d, tiles per depth\n", c->kernel,
                      c->es == 8 ? "fp64" : "fp32", Tmax);
  for (int T = 1; T <= Tmax && rc == GIRIH_OK; ++T) {
    double bestT = -1;
    int btile = 0;
    for (int tile : tile_candidates(c, T)) {
      c->opt_tile = tile;
      double ms = 0;
      rc = time_pass_local(c, T, 3, &ms);
      if (rc != GIRIH_OK) break;
      const double mlups = lups * T / ms / 1e3;
      if (verbose) printf("[AUTO TUNE]     [T:%d tile:%04d]  time:%e  MLUPS:%06llu\n", T, tile, ms * 1e-3, (unsigned long long)mlups);
      if (mlups > bestT) { bestT = mlups; btile = tile; }
    }
    c->tuned_tile[T & 7] = btile;
    if (bestT > best) { best = bestT; bT = T; }
  }
  c->opt_tile = saved_tile;
  if (rc == GIRIH_OK) {
    c->tuned_tfuse = bT;
    CU(cudaEventRecord(e1, c->s_comp));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (verbose) {
      printf("[AUTO TUNE] COMPLETE: fused steps per pass:%d  tile:%04d  perf:%7.2f MLUP/s\n", bT, c->tuned_tile[bT & 7], best);
      printf("[AUTO TUNE]  Tuning time: %5.3f seconds\n", ms * 1e-3);
      fflush(stdout);
    }
    if (best_tfuse) *best_tfuse = bT;
    if (best_tile) *best_tile = c->tuned_tile[bT & 7];
    if (best_mlups) *best_mlups = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

extern "C" int girih_gpu_last_elapsed_ms(girih_gpu_ctx *c, double *compute_ms, double *comm_ms, double *total_ms) {
  if (!c) return GIRIH_ERR_ARG;
  if (compute_ms) *compute_ms = c->ms_compute;
  if (comm_ms) *comm_ms = c->ms_comm;
  if (total_ms) *total_ms = c->ms_total;
  return GIRIH_OK;
}
extern "C" int girih_gpu_get_stat(girih_gpu_ctx *c, const char *key, long long *value) {
  if (!c || !key || !value) return GIRIH_ERR_ARG;
  if (!strcmp(key, "exact_launches")) *value = c->n_exact;
  else if (!strcmp(key, "fused_launches")) *value = c->n_fused;
  else return fail(c, GIRIH_ERR_ARG, "unknown stat '%s'", key);
  return GIRIH_OK;
}
extern "C" int girih_gpu_last_launch_info(girih_gpu_ctx *c, int *n_kernels, int *n_passes, int *n_steps, int *tfuse_used) {
  if (!c) return GIRIH_ERR_ARG;
  if (n_kernels) *n_kernels = c->n_kernels;
  if (n_passes) *n_passes = c->n_passes;
  if (n_steps) *n_steps = c->n_steps;
  if (tfuse_used) *tfuse_used = c->tfuse_used;
  return GIRIH_OK;
}

// ------------------------------------------------------------------------------------------------
// NaN / zero scan of U1 (src/utils.c:819-840)
// ------------------------------------------------------------------------------------------------
template <typename R>
__global__ void k_scan(DevGrid g, const R *__restrict__ u, int hx, int hy, int hz, unsigned long long *out) {
  // over the cells of the HOST array extent: (hx, hy, hz) starting at (X0-r, Y0-r, Z0-r)
  unsigned long long nans = 0, zeros = 0;
  const long long n = (long long)hx * hy * hz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % hx);
    const long long t = i / hx;
    const int y = (int)(t % hy), z = (int)(t / hy);
    const R v = u[((long long)(z + g.Z0 - g.r) * g.ny_dev + (y + g.Y0 - g.r)) * g.px + (x + g.X0 - g.r)];
    nans += (v * (R)0 != (R)0) ? 1 : 0;
    zeros += (fabs((double)v) < 1e-6) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    nans += __shfl_down_sync(0xffffffffu, nans, o);
    zeros += __shfl_down_sync(0xffffffffu, zeros, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nans) atomicAdd(out, nans);
    if (zeros) atomicAdd(out + 1, zeros);
  }
}

template <typename R>
__global__ void k_scan_flat(const R *__restrict__ u, long long n, unsigned long long *out) {
  unsigned long long nans = 0, zeros = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const R v = u[i];
    nans += (v * (R)0 != (R)0) ? 1 : 0;
    zeros += (fabs((double)v) < 1e-6) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    nans += __shfl_down_sync(0xffffffffu, nans, o);
    zeros += __shfl_down_sync(0xffffffffu, zeros, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nans) atomicAdd(out, nans);
    if (zeros) atomicAdd(out + 1, zeros);
  }
}

extern "C" int girih_gpu_scan_u1(girih_gpu_ctx *c, uint64_t *n_nan_inf, uint64_t *n_zero) {
  if (!c) return GIRIH_ERR_ARG;
  if (!c->uploaded) return fail(c, GIRIH_ERR_STATE, "scan before upload");
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(c->d_scan, 0, 2 * sizeof(unsigned long long), c->s_comp));
  if (c->solar) {
    // the reference scans the first ln_domain reals of the array (src/performance.c:131-141), i.e. the lower half of
    // the first field: the same reals here (host layout on the device)
    const long long n = c->solar_n2 / 2;
    if (c->es == 8) {
      auto kfn = k_scan_flat<double>;
      GIRIH_LAUNCH(kfn, 148 * 8, 256, 0, c->s_comp, (const double *)c->dU[0], n, c->d_scan);
    } else {
      auto kfn = k_scan_flat<float>;
      GIRIH_LAUNCH(kfn, 148 * 8, 256, 0, c->s_comp, (const float *)c->dU[0], n, c->d_scan);
    }
    CU(cudaGetLastError());
    unsigned long long hs[2];
    CU(cudaMemcpyAsync(hs, c->d_scan, sizeof(hs), cudaMemcpyDeviceToHost, c->s_comp));
    CU(cudaStreamSynchronize(c->s_comp));
    if (n_nan_inf) *n_nan_inf = hs[0];
    if (n_zero) *n_zero = hs[1];
    return GIRIH_OK;
  }
  // the reference scans all ln_domain cells including the x padding, which is zero; count the
  // padding cells as zeros to report the same percentage
  const int hx = c->g.nx + 2 * c->g.r;
  if (c->es == 8) {
    auto kfn = k_scan<double>;
    GIRIH_LAUNCH(kfn, 148 * 8, 256, 0, c->s_comp, c->g, (const double *)c->dU[0], hx, c->hshape[1], c->hshape[2], c->d_scan);
  } else {
    auto kfn = k_scan<float>;
    GIRIH_LAUNCH(kfn, 148 * 8, 256, 0, c->s_comp, c->g, (const float *)c->dU[0], hx, c->hshape[1], c->hshape[2], c->d_scan);
  }
  CU(cudaGetLastError());
  unsigned long long h[2];
  CU(cudaMemcpyAsync(h, c->d_scan, sizeof(h), cudaMemcpyDeviceToHost, c->s_comp));
  CU(cudaStreamSynchronize(c->s_comp));
  const unsigned long long pad = (unsigned long long)(c->hshape[0] - hx) * c->hshape[1] * c->hshape[2];
  if (n_nan_inf) *n_nan_inf = h[0];
  if (n_zero) *n_zero = h[1] + pad;
  return GIRIH_OK;
}
