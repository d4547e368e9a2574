// emu_driver.cpp -- C entry points (ctypes) that run the library's own launchers (launch_r1, launch_r4,
// launch_box from girih_b200/csrc/inst_*.cu, compiled for the emulator) plus k_naive on host memory laid
// out exactly like the device arrays (layout.h).  TEST INFRASTRUCTURE; see include/cuda_runtime.h.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "kernels_naive.cuh"
#include "launch.h"
#include "layout.h"
#include "stencil_expr.cuh"

using namespace girih;

namespace {

struct EmuCtx {
  int kernel, es, r, time_order, ncoef;
  int hshape[3];
  DevGrid g;
  size_t arr_elems;
  void *dU[2] = {nullptr, nullptr};
  void *dU3 = nullptr, *dCoef = nullptr;
  double cc[5] = {0, 0, 0, 0, 0};
  int launches = 0;
};

void *zalloc(size_t bytes) {
  void *p = nullptr;
  if (posix_memalign(&p, 256, bytes ? bytes : 256) != 0) return nullptr;
  memset(p, 0, bytes);
  return p;
}

// host array (reference layout [z][y][x], shape hshape) <-> DevGrid layout
template <typename R> void repitch(const EmuCtx *c, R *dev, R *host, bool to_dev) {
  const DevGrid &g = c->g;
  const int r = c->r;
  for (int k = 0; k < c->hshape[2]; ++k)
    for (int j = 0; j < c->hshape[1]; ++j) {
      R *h = host + ((size_t)k * c->hshape[1] + j) * c->hshape[0];
      R *d = dev + ((long long)(k - r + g.Z0) * g.ny_dev + (j - r + g.Y0)) * g.px + (g.X0 - r);
      // the host row may carry alignment padding beyond nx + 2r (src/utils.c:367-374): it is never read
      const int n = g.nx + 2 * r;
      if (to_dev) memcpy(d, h, sizeof(R) * (size_t)n);
      else memcpy(h, d, sizeof(R) * (size_t)n);
    }
}

template <int K, typename R, bool FM>
cudaError_t naive_t(EmuCtx *c, int dst) {
  const DevGrid &g = c->g;
  dim3 block(64, 4, 1);
  dim3 grid((g.nx + 63) / 64, (g.ny + 3) / 4, g.nz);
  ConstCoef<R> k;
  for (int i = 0; i < 5; ++i) k.v[i] = (R)c->cc[i];
  auto kfn = k_naive<K, R, FM>;
  GIRIH_LAUNCH(kfn, grid, block, 0, nullptr, g, (R *)c->dU[dst], (const R *)c->dU[dst ^ 1], (const R *)c->dU3,
               (const R *)c->dCoef, (long long)c->arr_elems, k, g.X0, g.Y0, g.Z0, g.X0 + g.nx, g.Y0 + g.ny,
               g.Z0 + g.nz);
  return cudaGetLastError();
}

template <bool FM> cudaError_t naive(EmuCtx *c, int dst) {
#define GN(K) case K: return c->es == 8 ? naive_t<K, double, FM>(c, dst) : naive_t<K, float, FM>(c, dst);
  switch (c->kernel) { GN(0) GN(1) GN(2) GN(3) GN(4) GN(5) GN(7) default: return cudaErrorInvalidValue; }
#undef GN
}

}  // namespace

#define EMU_API __attribute__((visibility("default")))
extern "C" {

// r, time_order, number of per-point coefficient arrays and max_tfuse as the library's kernel table has them
// (passed in by the test from girih_kernel_info so the two cannot drift apart)
EMU_API void *emu_create(int kernel, int es, const int st[3], const int hshape[3], int r, int time_order, int ncoef,
                 int max_tfuse) {
  EmuCtx *c = new EmuCtx();
  c->kernel = kernel; c->es = es; c->r = r; c->time_order = time_order; c->ncoef = ncoef;
  for (int d = 0; d < 3; ++d) c->hshape[d] = hshape[d];
  make_dev_grid(c->g, st, r, max_tfuse, es, 0, 1);
  c->arr_elems = (size_t)c->g.pxy * c->g.nz_dev;
  const size_t bytes = c->arr_elems * (size_t)es;
  c->dU[0] = zalloc(bytes);
  c->dU[1] = zalloc(bytes);
  if (time_order == 2) c->dU3 = zalloc(bytes);
  if (ncoef > 0) c->dCoef = zalloc(bytes * (size_t)ncoef);
  return c;
}

EMU_API void emu_destroy(void *p) {
  EmuCtx *c = (EmuCtx *)p;
  if (!c) return;
  free(c->dU[0]); free(c->dU[1]); free(c->dU3); free(c->dCoef);
  delete c;
}

// U1 -> array 0, U2 -> array 1 (girih_gpu_upload's convention); coef: scalar coefficients (ncoef == 0,
// first 5 values used) or the per-point arrays, array m at host offset m * (hshape product)
EMU_API int emu_upload(void *p, void *U1, void *U2, void *U3, void *coef, int n_scalar) {
  EmuCtx *c = (EmuCtx *)p;
  const size_t hn = (size_t)c->hshape[0] * c->hshape[1] * c->hshape[2];
  if (c->es == 8) {
    repitch<double>(c, (double *)c->dU[0], (double *)U1, true);
    repitch<double>(c, (double *)c->dU[1], (double *)U2, true);
    if (c->dU3) repitch<double>(c, (double *)c->dU3, (double *)U3, true);
    for (int m = 0; m < c->ncoef; ++m)
      repitch<double>(c, (double *)c->dCoef + (size_t)m * c->arr_elems, (double *)coef + (size_t)m * hn, true);
    if (c->ncoef == 0) for (int i = 0; i < 5 && i < n_scalar; ++i) c->cc[i] = ((double *)coef)[i];
  } else {
    repitch<float>(c, (float *)c->dU[0], (float *)U1, true);
    repitch<float>(c, (float *)c->dU[1], (float *)U2, true);
    if (c->dU3) repitch<float>(c, (float *)c->dU3, (float *)U3, true);
    for (int m = 0; m < c->ncoef; ++m)
      repitch<float>(c, (float *)c->dCoef + (size_t)m * c->arr_elems, (float *)coef + (size_t)m * hn, true);
    if (c->ncoef == 0) for (int i = 0; i < 5 && i < n_scalar; ++i) c->cc[i] = (double)((float *)coef)[i];
  }
  return 0;
}

EMU_API int emu_download(void *p, void *U1, void *U2) {
  EmuCtx *c = (EmuCtx *)p;
  if (c->es == 8) {
    repitch<double>(c, (double *)c->dU[0], (double *)U1, false);
    repitch<double>(c, (double *)c->dU[1], (double *)U2, false);
  } else {
    repitch<float>(c, (float *)c->dU[0], (float *)U1, false);
    repitch<float>(c, (float *)c->dU[1], (float *)U2, false);
  }
  return 0;
}

// One pass of T fused steps reading array `src`, writing array `src ^ 1`, over output planes
// [zb, ze) (local interior planes, 0-based) and optionally a second range [zb1, ze1); the arguments
// mirror StreamLaunch.  variant: 0 auto, 1 = k_naive, 2 = fused-sweep kernel also for T = 1.
EMU_API int emu_pass(void *p, int T, int src, int zb, int ze, int zb1, int ze1, int tile, int variant, int contract,
             int zchunk) {
  EmuCtx *c = (EmuCtx *)p;
  const DevGrid &g = c->g;
  c->launches++;
  if (variant == 1) {
    if (T != 1) return cudaErrorInvalidValue;
    return contract ? naive<true>(c, src ^ 1) : naive<false>(c, src ^ 1);
  }
  StreamLaunch sl;
  sl.g = g;
  sl.in = c->dU[src];
  sl.out = c->dU[src ^ 1];
  sl.roc2 = c->dU3;
  sl.coef = c->dCoef;
  sl.coef_stride = (long long)c->arr_elems;
  for (int i = 0; i < 5; ++i) sl.cc[i] = c->cc[i];
  sl.zb0 = g.Z0 + zb;
  sl.ze0 = g.Z0 + ze;
  sl.zb1 = ze1 > zb1 ? g.Z0 + zb1 : 0;
  sl.ze1 = ze1 > zb1 ? g.Z0 + ze1 : 0;
  sl.zchunk = zchunk;
  sl.tile = tile;
  sl.variant = variant;
  sl.contract = contract;
  sl.stream = nullptr;
  if (c->kernel == 7) return T == 1 ? launch_box(c->es, sl) : cudaErrorInvalidValue;
  if (g.r == 1) return launch_r1(c->kernel, c->es, T, sl);
  if (T != 1) return cudaErrorInvalidValue;
  return launch_r4(c->kernel, c->es, sl);
}

}  // extern "C"
