// instantiation unit: 27-point box operator (slot 7), both precisions
#include <algorithm>

#include "kernels_box.cuh"
#include "launch.h"

namespace girih {

template <typename R, int NW, bool FM>
static cudaError_t launch_box_t(const StreamLaunch &s) {
  const DevGrid &g = s.g;
  constexpr int WX = 32 * Vec<R>::N;
  BoxArgs<R> a;
  a.g = g;
  a.in = (const R *)s.in;
  a.out = (R *)s.out;
  for (int i = 0; i < 5; ++i) a.cc.v[i] = (R)s.cc[i];
  a.zb0 = s.zb0;
  a.ze0 = s.ze0;
  const int nz = s.ze0 - s.zb0;
  const int ntiles = ((g.nx + WX - 1) / WX) * ((g.ny + NW - 1) / NW);
  int zchunk = s.zchunk > 0 ? s.zchunk : 64;
  if (s.zchunk <= 0) {   // every chunk re-reads two planes; keep chunks long but leave several waves of CTAs
    while (zchunk < nz && (long long)ntiles * ((nz + zchunk - 1) / zchunk) > 148LL * 8 * 8) zchunk *= 2;
    while (zchunk > 16 && (long long)ntiles * ((nz + zchunk - 1) / zchunk) < 148LL * 8 * 2) zchunk /= 2;
  }
  zchunk = std::min(zchunk, std::max(nz, 1));
  a.zchunk = zchunk;
  dim3 grid((g.nx + WX - 1) / WX, (g.ny + NW - 1) / NW, (nz + zchunk - 1) / zchunk);
  auto kfn = k_box_march<R, NW, FM>;
  GIRIH_LAUNCH(kfn, grid, 32 * NW, 0, s.stream, a);
  return cudaGetLastError();
}

template <typename R, int NW, bool FM>
static cudaError_t launch_box_async_t(const StreamLaunch &s) {
  using Cfg = BoxACfg<R, NW>;
  const DevGrid &g = s.g;
  BoxArgs<R> a;
  a.g = g;
  a.in = (const R *)s.in;
  a.out = (R *)s.out;
  for (int i = 0; i < 5; ++i) a.cc.v[i] = (R)s.cc[i];
  a.zb0 = s.zb0;
  a.ze0 = s.ze0;
  const int nz = s.ze0 - s.zb0;
  int zchunk = s.zchunk > 0 ? s.zchunk : 64;
  zchunk = std::min(zchunk, std::max(nz, 1));
  a.zchunk = zchunk;
  dim3 grid((g.nx + Cfg::WX - 1) / Cfg::WX, (g.ny + NW - 1) / NW, (nz + zchunk - 1) / zchunk);
  auto kfn = k_box_async<R, NW, FM>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  if (e != cudaSuccess) return e;
  GIRIH_LAUNCH(kfn, grid, 32 * NW, Cfg::SMEM, s.stream, a);
  return cudaGetLastError();
}

template <typename R, int NW, bool FM>
static cudaError_t launch_box_lean_t(const StreamLaunch &s) {
  using Cfg = BoxACfg<R, NW>;
  const DevGrid &g = s.g;
  BoxArgs<R> a;
  a.g = g;
  a.in = (const R *)s.in;
  a.out = (R *)s.out;
  for (int i = 0; i < 5; ++i) a.cc.v[i] = (R)s.cc[i];
  a.zb0 = s.zb0;
  a.ze0 = s.ze0;
  const int nz = s.ze0 - s.zb0;
  int zchunk = s.zchunk > 0 ? s.zchunk : 64;
  zchunk = std::min(zchunk, std::max(nz, 1));
  a.zchunk = zchunk;
  dim3 grid((g.nx + Cfg::WX - 1) / Cfg::WX, (g.ny + NW - 1) / NW, (nz + zchunk - 1) / zchunk);
  auto kfn = k_box_lean<R, NW, FM>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  if (e != cudaSuccess) return e;
  GIRIH_LAUNCH(kfn, grid, 32 * NW, Cfg::SMEM, s.stream, a);
  return cudaGetLastError();
}

// tile option: 4 / 8 = the register-marching kernel with that many rows per CTA, 108 / 116 = the cp.async kernel with 8 / 16
// rows, 208 / 216 = its lean form (running pointers, predicated copies: 35% fewer non-arithmetic instructions);
// default = cp.async, 8 rows (round 2, 512^3 sustained: fp64 240 against 164 GLUP/s, fp32 492-511 against 397-412;
// profiles/r02_box_async.log)
cudaError_t launch_box(int es, const StreamLaunch &s) {
  if (s.contract) return es == 8 ? launch_box_async_t<double, 8, true>(s) : launch_box_async_t<float, 8, true>(s);
  if (s.tile == 4) return es == 8 ? launch_box_t<double, 4, false>(s) : launch_box_t<float, 4, false>(s);
  if (s.tile == 8) return es == 8 ? launch_box_t<double, 8, false>(s) : launch_box_t<float, 8, false>(s);
  if (s.tile == 108) return es == 8 ? launch_box_async_t<double, 8, false>(s) : launch_box_async_t<float, 8, false>(s);
  if (s.tile == 116) return es == 8 ? launch_box_async_t<double, 16, false>(s) : launch_box_async_t<float, 16, false>(s);
  if (s.tile == 208) return es == 8 ? launch_box_lean_t<double, 8, false>(s) : launch_box_lean_t<float, 8, false>(s);
  if (s.tile == 216) return es == 8 ? launch_box_lean_t<double, 16, false>(s) : launch_box_lean_t<float, 16, false>(s);
  // default: fp64 = cp.async ring (the FP64 pipe and its latencies bound it: 256 against 245 GLUP/s for the lean form),
  // fp32 = the lean form of the same ring (535 against 487; profiles/r02_box_lean.log, 512^3 sustained)
  return es == 8 ? launch_box_async_t<double, 8, false>(s) : launch_box_lean_t<float, 8, false>(s);
}

}  // namespace girih
