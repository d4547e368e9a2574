// instantiation unit: radius-4 operators (slots 0 and 4), both precisions
#include <algorithm>

#include "kernels_r4.cuh"
#include "kernels_r4_strip.cuh"
#include "launch.h"

namespace girih {

template <int K, typename R, int NW>
static cudaError_t launch_r4_t(const StreamLaunch &s) {
  using Cfg = R4Cfg<R, NW>;
  const DevGrid &g = s.g;
  R4Args<R> a;
  a.g = g;
  a.v = (const R *)s.in;
  a.u = (R *)s.out;
  a.roc2 = (const R *)s.roc2;
  a.coef = (const R *)s.coef;
  a.coef_stride = s.coef_stride;
  for (int i = 0; i < 5; ++i) a.cc.v[i] = (R)s.cc[i];
  a.zb0 = s.zb0;
  a.ze0 = s.ze0;
  const int ntx = (g.nx + Cfg::WX - 1) / Cfg::WX, nty = (g.ny + Cfg::H - 1) / Cfg::H;
  int zchunk = s.zchunk;
  if (zchunk <= 0) zchunk = std::max(1, std::min(s.ze0 - s.zb0, 32));   // measured best on B200 (kernel sweep)
  a.zchunk = zchunk;
  dim3 grid(ntx, nty, (s.ze0 - s.zb0 + zchunk - 1) / zchunk);
  auto kfn = k_r4<K, R, NW>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  if (e != cudaSuccess) return e;
  GIRIH_LAUNCH(kfn, grid, 32 * NW, Cfg::SMEM, s.stream, a);
  return cudaGetLastError();
}

template <int K, typename R, int PY, int NW, bool FM = false>
static cudaError_t launch_r4_strip_t(const StreamLaunch &s) {
  using Cfg = R4StripCfg<R, PY, NW>;
  const DevGrid &g = s.g;
  R4Args<R> a;
  a.g = g;
  a.v = (const R *)s.in;
  a.u = (R *)s.out;
  a.roc2 = (const R *)s.roc2;
  a.coef = (const R *)s.coef;
  a.coef_stride = s.coef_stride;
  for (int i = 0; i < 5; ++i) a.cc.v[i] = (R)s.cc[i];
  a.zb0 = s.zb0;
  a.ze0 = s.ze0;
  const int ntx = (g.nx + Cfg::WX - 1) / Cfg::WX, nty = (g.ny + Cfg::H - 1) / Cfg::H;
  int zchunk = s.zchunk;
  if (zchunk <= 0) {
    const int nz = s.ze0 - s.zb0;
    const int want = 148 * 8;
    const int nch = std::max(1, (want + ntx * nty - 1) / (ntx * nty));
    zchunk = std::max(std::min(nz, 32), (nz + nch - 1) / nch);
  }
  a.zchunk = zchunk;
  dim3 grid(ntx, nty, (s.ze0 - s.zb0 + zchunk - 1) / zchunk);
  auto kfn = k_r4_strip<K, R, PY, NW, FM>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  if (e != cudaSuccess) return e;
  GIRIH_LAUNCH(kfn, grid, 32 * NW, Cfg::SMEM, s.stream, a);
  return cudaGetLastError();
}

template <int K, typename R, int NW, bool FM = false>
static cudaError_t launch_r4_async_t(const StreamLaunch &s) {
  using Cfg = R4Cfg<R, NW>;
  using ACfg = R4ACfg<R, NW>;
  const DevGrid &g = s.g;
  R4Args<R> a;
  a.g = g;
  a.v = (const R *)s.in;
  a.u = (R *)s.out;
  a.roc2 = (const R *)s.roc2;
  a.coef = (const R *)s.coef;
  a.coef_stride = s.coef_stride;
  for (int i = 0; i < 5; ++i) a.cc.v[i] = (R)s.cc[i];
  a.zb0 = s.zb0;
  a.ze0 = s.ze0;
  const int ntx = (g.nx + Cfg::WX - 1) / Cfg::WX, nty = (g.ny + Cfg::H - 1) / Cfg::H;
  int zchunk = s.zchunk;
  if (zchunk <= 0) zchunk = std::max(1, std::min(s.ze0 - s.zb0, 64));
  a.zchunk = zchunk;
  dim3 grid(ntx, nty, (s.ze0 - s.zb0 + zchunk - 1) / zchunk);
  auto kfn = k_r4_async<K, R, NW, FM>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ACfg::SMEM);
  if (e != cudaSuccess) return e;
  GIRIH_LAUNCH(kfn, grid, 32 * NW, ACfg::SMEM, s.stream, a);
  return cudaGetLastError();
}

cudaError_t launch_r4(int kernel, int es, const StreamLaunch &s) {
  // tile option = warps (rows) per CTA
  // tile option: 8 / 16 = ring variant with that many rows per CTA, 108 / 116 = cp.async variant with 8 / 16 rows;
  // default = cp.async variant with 16 rows per CTA: 1-2% ahead of 8 rows once the power controller has settled under
  // the 1 000 W cap (768^3: fp64 177.8 against 175.0 GLUP/s, fp32 347.5 against 344-346; profiles/r02_k0_sustained2.log).
  // The 2-rows-per-thread strip kernel of slot 4 was tried for slot 0: 114 / 234 GLUP/s.
  if (s.contract) {   // contracted arithmetic: default kernels only
    if (kernel == 0) return es == 8 ? launch_r4_async_t<0, double, 16, true>(s) : launch_r4_async_t<0, float, 16, true>(s);
    if (kernel == 4) return es == 8 ? launch_r4_strip_t<4, double, 2, 8, true>(s) : launch_r4_strip_t<4, float, 2, 8, true>(s);
    return cudaErrorInvalidValue;
  }
  if (kernel == 0) {
    if (s.tile == 8) return es == 8 ? launch_r4_t<0, double, 8>(s) : launch_r4_t<0, float, 8>(s);
    if (s.tile == 16) return es == 8 ? launch_r4_t<0, double, 16>(s) : launch_r4_t<0, float, 16>(s);
    if (s.tile == 116) return es == 8 ? launch_r4_async_t<0, double, 16>(s) : launch_r4_async_t<0, float, 16>(s);
    if (s.tile == 108) return es == 8 ? launch_r4_async_t<0, double, 8>(s) : launch_r4_async_t<0, float, 8>(s);
    // thin z ranges (the outer parts of a slab in the halo-first schedule, slabs of a strong-scaling run on many GPUs)
    // give the 16-row tile too few CTAs: 8 rows there (C5 on 8 GPUs, 128 planes per GPU: 2 526 against 2 336 GLUP/s)
    if (s.ze0 - s.zb0 < 96) return es == 8 ? launch_r4_async_t<0, double, 8>(s) : launch_r4_async_t<0, float, 8>(s);
    return es == 8 ? launch_r4_async_t<0, double, 16>(s) : launch_r4_async_t<0, float, 16>(s);
  }
  if (kernel == 4) {
    if (s.tile == 16) return es == 8 ? launch_r4_t<4, double, 16>(s) : launch_r4_t<4, float, 16>(s);
    if (s.tile == 8) return es == 8 ? launch_r4_t<4, double, 8>(s) : launch_r4_t<4, float, 8>(s);
    return es == 8 ? launch_r4_strip_t<4, double, 2, 8>(s) : launch_r4_strip_t<4, float, 2, 8>(s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace girih
