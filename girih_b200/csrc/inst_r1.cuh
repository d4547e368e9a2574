// inst_r1.cuh -- host-side launcher of k_r1 for one (operator, precision); included by the
// inst_r1_k*_f*.cu translation units.
#pragma once
#include <algorithm>
#ifdef GIRIH_R1X_TRACE
#include <stdio.h>
#include <stdlib.h>

#include <vector>
#endif

#include "kernels_r1.cuh"
#include "kernels_r1_march.cuh"
#include "kernels_r1x.cuh"
#include "launch.h"

namespace girih {

template <typename R> static inline ConstCoef<R> make_cc(const double (&cc)[5]) {
  ConstCoef<R> k;
  for (int i = 0; i < 5; ++i) k.v[i] = (R)cc[i];
  return k;
}

// Per kernel instantiation and device: has the dynamic shared-memory limit been raised, how many CTAs fit an SM, how many
// SMs are there.  Asked once instead of on every launch (the queries sat on the launch path of every pass).
// Rank threads of one process may race on an entry; they all write the same values.
template <int K, typename R, int T, int PY, int NW, int DBG> struct R1LaunchCache {
  static inline bool attr_done[64] = {};
  static inline int occ[64] = {};
  static inline int nsm[64] = {};
};

template <int K, typename R, int T, int PY, int NW, int DBG = 0>
static cudaError_t launch_r1_t(const StreamLaunch &s) {
  using Cfg = R1Cfg<R, T, PY, NW>;
  using Cache = R1LaunchCache<K, R, T, PY, NW, DBG>;
  const DevGrid &g = s.g;
  R1Args<R> a;
  a.g = g;
  a.in = (const R *)s.in;
  a.out = (R *)s.out;
  a.coef = (const R *)s.coef;
  a.coef_stride = s.coef_stride;
  a.cc = make_cc<R>(s.cc);
  a.zb0 = s.zb0;
  a.ze0 = s.ze0;
  a.push_up = (R *)s.push_up;
  a.push_dn = (R *)s.push_dn;
  a.push_up_from = s.push_up ? s.push_up_from : 0x7fffffff;
  a.push_dn_below = s.push_dn ? s.push_dn_below : -0x7fffffff;
  const int ntx = (g.nx + Cfg::UX - 1) / Cfg::UX, nty = (g.ny + Cfg::UY - 1) / Cfg::UY;
  auto kfn = k_r1<K, R, T, PY, NW, DBG>;
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev & 63;
  if (!Cache::attr_done[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    int occ = 1, nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, 32 * NW, Cfg::SMEM) != cudaSuccess || occ < 1) occ = 1;
    Cache::occ[slot] = occ;
    Cache::nsm[slot] = nsm;
    Cache::attr_done[slot] = true;
  }
  int zchunk = s.zchunk;
  if (zchunk <= 0) {
    // Split z into chunks so that the CTAs fill whole waves of the 148 SMs: every chunk pays 2T planes
    // of pipeline fill, every partially filled last wave idles SMs.  Minimise
    //   waves(ntiles * nch) * (ceil(nz / nch) + 2T).
    const int occ = Cache::occ[slot], nsm = Cache::nsm[slot];
    // a second range of the same length doubles the CTAs per chunk row
    const int nz = s.ze0 - s.zb0, ntiles = ntx * nty * (s.ze1 > s.zb1 ? 2 : 1), slots = nsm * occ;
    long long best = -1;
    zchunk = nz;
    for (int nch = 1; nch <= nz && nch <= 64; ++nch) {
      const int zc = (nz + nch - 1) / nch;
      if (zc < 4 * T && nch > 1) break;
      const int real_nch = (nz + zc - 1) / zc;
      const long long waves = ((long long)ntiles * real_nch + slots - 1) / slots;
      const long long cost = waves * (zc + 2 * T);
      if (best < 0 || cost < best) { best = cost; zchunk = zc; }
    }
  }
  a.zchunk = zchunk;
  a.nch0 = (s.ze0 - s.zb0 + zchunk - 1) / zchunk;
  a.zb1 = s.zb1;
  a.ze1 = s.ze1;
  const int nch1 = (s.ze1 > s.zb1) ? (s.ze1 - s.zb1 + zchunk - 1) / zchunk : 0;
  dim3 grid(ntx, nty, a.nch0 + nch1);
  GIRIH_LAUNCH(kfn, grid, 32 * NW, Cfg::SMEM, s.stream, a);
  return cudaGetLastError();
}

// Exact (non-overlapping) tiles with edge hand-off between co-resident CTAs (kernels_r1x.cuh).  Needs one resident CTA
// per tile and z chunk: cudaErrorNotSupported tells the caller to take the overlapped-tile kernel instead.
template <int K, typename R, int T, int PY, int NW, bool FM = false, int XDBG = 0>
static cudaError_t launch_r1x_t(const StreamLaunch &s) {
  if constexpr (sizeof(R) != 8 || T < 3) {
    return cudaErrorNotSupported;
  } else {
    using Cfg = R1xCfg<R, T, PY, NW>;
    const DevGrid &g = s.g;
    if (s.xbuf == nullptr || s.xseq == nullptr || s.ze1 > s.zb1 || s.push_up != nullptr || s.push_dn != nullptr)
      return cudaErrorNotSupported;
    const int ntx = (g.nx + Cfg::WX - 1) / Cfg::WX, nty = (g.ny + Cfg::H - 1) / Cfg::H, nz = s.ze0 - s.zb0;
    const long long tiles = (long long)ntx * nty;
    if (nz < 1 || tiles > s.nsm) return cudaErrorNotSupported;
    // z chunks: as many as stay co-resident, each at least 4T planes long (a chunk pays 2T planes of pipeline fill)
    int nch = (int)std::min<long long>(s.nsm / tiles, std::max(1, nz / (4 * T)));
    if (s.zchunk > 0) nch = std::min(nch, std::max(1, (nz + s.zchunk - 1) / s.zchunk));
    const int zchunk = (nz + nch - 1) / nch;
    nch = (nz + zchunk - 1) / zchunk;
    if ((size_t)(tiles * nch) * Cfg::TILE_BYTES > s.xbuf_bytes) return cudaErrorNotSupported;
    R1xArgs<R> a;
    a.g = g;
    a.in = (const R *)s.in;
    a.out = (R *)s.out;
    a.coef = (const R *)s.coef;
    a.coef_stride = s.coef_stride;
    a.cc = make_cc<R>(s.cc);
    a.zb0 = s.zb0;
    a.ze0 = s.ze0;
    a.zchunk = zchunk;
    a.xbuf = s.xbuf;
    a.err = s.xerr;
    a.seq0 = *s.xseq;
    auto kfn = k_r1x<K, R, T, PY, NW, FM, XDBG>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    dim3 grid(ntx, nty, nch);
#ifdef GIRIH_R1X_TRACE
    static long long *d_trace = nullptr;
    const size_t trace_n = (size_t)tiles * nch * 2 * 16 * T * 2;
    if (!d_trace) cudaMalloc((void **)&d_trace, 148 * 2 * 16 * 8 * 2 * sizeof(long long));
    cudaMemsetAsync(d_trace, 0, trace_n * sizeof(long long), s.stream);
    a.trace = d_trace;
#endif
    e = GIRIH_LAUNCH_COOP(kfn, grid, dim3(32 * NW), Cfg::SMEM, s.stream, a);
#ifdef GIRIH_R1X_TRACE
    if (const char *path = getenv("GIRIH_R1X_TRACE_FILE")) {
      cudaStreamSynchronize(s.stream);
      std::vector<long long> h(trace_n);
      cudaMemcpy(h.data(), d_trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost);
      if (FILE *f = fopen(path, "wb")) {
        const int hdr[4] = {ntx, nty, nch, T};
        fwrite(hdr, sizeof(int), 4, f);
        fwrite(h.data(), sizeof(long long), trace_n, f);
        fclose(f);
      }
    }
#endif
    if (e == cudaErrorCooperativeLaunchTooLarge) {   // something else holds SMs: overlapped tiles need no co-residency
      (void)cudaGetLastError();
      return cudaErrorNotSupported;
    }
    if (e == cudaSuccess) *s.xseq += (unsigned)(zchunk + 2 * T) + 2u;   // tags of this launch: seq0 + 1 .. seq0 + nit
    return e;
  }
}

// tile shapes instantiated per (operator, precision, depth); "tile" option = PY*100 + NW
template <int K, typename R, int T>
static cudaError_t launch_r1_tile(const StreamLaunch &s) {
  if constexpr (K == 1 && sizeof(R) == 8 && T >= 3) {
    // exact tiles: 10000 + PY*100 + NW (slot 1, fp64); falls through to the overlapped tiles when the grid cannot be co-resident
    if (tile_is_exact(s.tile)) {
      cudaError_t e = cudaErrorNotSupported;
      if (s.tile == 10408) e = s.contract ? launch_r1x_t<K, R, T, 4, 8, true>(s) : launch_r1x_t<K, R, T, 4, 8, false>(s);
      if (s.tile == 10216) e = s.contract ? launch_r1x_t<K, R, T, 2, 16, true>(s) : launch_r1x_t<K, R, T, 2, 16, false>(s);
#ifdef GIRIH_PERF_EXPERIMENTS
      if constexpr (T == 4) {   // parts of the exchange removed (results INVALID): what each part costs per iteration
        if (s.tile == 11408) e = launch_r1x_t<K, R, T, 4, 8, false, 1>(s);
        if (s.tile == 12408) e = launch_r1x_t<K, R, T, 4, 8, false, 2>(s);
        if (s.tile == 13408) e = launch_r1x_t<K, R, T, 4, 8, false, 3>(s);
        if (s.tile == 17408) e = launch_r1x_t<K, R, T, 4, 8, false, 7>(s);
        if (s.tile == 13216) e = launch_r1x_t<K, R, T, 2, 16, false, 3>(s);
        if (s.tile == 17216) e = launch_r1x_t<K, R, T, 2, 16, false, 7>(s);
        if (s.tile == 18408) e = launch_r1x_t<K, R, T, 4, 8, false, 11>(s);   // rim values parked in shared memory only
        if (s.tile == 18216) e = launch_r1x_t<K, R, T, 2, 16, false, 11>(s);
      }
#endif
      if (e != cudaErrorNotSupported) return e;
    }
  }
  if constexpr (K == 1) {   // halo push: default tile of slot 1 only (girih_cuda.cu requests it for nothing else)
    if (s.push_up != nullptr || s.push_dn != nullptr) {
      if constexpr (sizeof(R) == 8) return s.contract ? launch_r1_t<K, R, T, 4, 8, R1_FM | R1_PUSH>(s) : launch_r1_t<K, R, T, 4, 8, R1_PUSH>(s);
      else return s.contract ? launch_r1_t<K, R, T, 2, 16, R1_FM | R1_PUSH>(s) : launch_r1_t<K, R, T, 2, 16, R1_PUSH>(s);
    }
  }
  if (s.contract) {   // contracted arithmetic: the default tile of each (operator, precision) only
    if constexpr (K == 1) {
      if (s.tile == 5408) return launch_r1_t<K, R, T, 4, 8, R1_FM | R1_SPLIT>(s);
      if (s.tile == 5216) return launch_r1_t<K, R, T, 2, 16, R1_FM | R1_SPLIT>(s);
      if (s.tile == 7408) return launch_r1_t<K, R, T, 4, 8, R1_FM | R1_REV>(s);
      if (s.tile == 7216) return launch_r1_t<K, R, T, 2, 16, R1_FM | R1_REV>(s);
      if (s.tile == 9408) return launch_r1_t<K, R, T, 4, 8, R1_FM | (T >= 4 ? R1_TRAP : 0)>(s);
      if (s.tile == 9216) return launch_r1_t<K, R, T, 2, 16, R1_FM | (T >= 2 ? R1_TRAP : 0)>(s);
    }
    if constexpr (KTraits<K>::NCA == 0 && sizeof(R) == 8) return launch_r1_t<K, R, T, 4, 8, R1_FM>(s);
    else return launch_r1_t<K, R, T, 2, 16, R1_FM>(s);
  }
  if (s.tile == 216) return launch_r1_t<K, R, T, 2, 16>(s);
  if (s.tile == 408) return launch_r1_t<K, R, T, 4, 8>(s);
#ifdef GIRIH_PERF_EXPERIMENTS
  // timing experiments of profiles/kernel_sweep_r01.md (parts of the kernel removed: results INVALID).
  // Not compiled into the product: build with `make lib PTXASV=-DGIRIH_PERF_EXPERIMENTS` to reproduce them.
  if constexpr (K == 1 && sizeof(R) == 8 && T == 4) {
    if (s.tile == 1408) return launch_r1_t<K, R, T, 4, 8, 1>(s);
    if (s.tile == 8408) return launch_r1_t<K, R, T, 4, 8, 8>(s);
    if (s.tile == 16408) return launch_r1_t<K, R, T, 4, 8, 16>(s);
    if (s.tile == 2408) return launch_r1_t<K, R, T, 4, 8, 2>(s);
    if (s.tile == 3408) return launch_r1_t<K, R, T, 4, 8, 3>(s);
    if (s.tile == 3216) return launch_r1_t<K, R, T, 2, 16, 3>(s);
    if (s.tile == 3312) return launch_r1_t<K, R, T, 3, 12, 3>(s);
    if (s.tile == 1216) return launch_r1_t<K, R, T, 2, 16, 1>(s);
    if (s.tile == 2216) return launch_r1_t<K, R, T, 2, 16, 2>(s);
  }
#endif
  // trapezoid skip (kernels_r1.cuh, TRAP): 9000 + tile, every radius-1 operator (the per-point coefficient loads of
  // the skipped level go as well).  A warp can only lie outside the core when PY <= T: below that depth the tile is
  // the plain one (same instantiation, no second copy of the kernel)
  if (s.tile == 9408) return launch_r1_t<K, R, T, 4, 8, (T >= 4 ? R1_TRAP : 0)>(s);
  if (s.tile == 9216) return launch_r1_t<K, R, T, 2, 16, (T >= 2 ? R1_TRAP : 0)>(s);
  if constexpr (K == 1) {
    // split-barrier variants: 5000 + tile
    if (s.tile == 5408) return launch_r1_t<K, R, T, 4, 8, R1_SPLIT>(s);
    if (s.tile == 5216) return launch_r1_t<K, R, T, 2, 16, R1_SPLIT>(s);
    // decoupled levels (kernels_r1.cuh, REV): 7000 + tile
    if (s.tile == 7408) return launch_r1_t<K, R, T, 4, 8, R1_REV>(s);
    if (s.tile == 7216) return launch_r1_t<K, R, T, 2, 16, R1_REV>(s);
    if (s.tile == 312) return launch_r1_t<K, R, T, 3, 12>(s);
    if (s.tile == 310) return launch_r1_t<K, R, T, 3, 10>(s);
    if (s.tile == 316) return launch_r1_t<K, R, T, 3, 16>(s);
  }
  // defaults from the B200 sweeps (profiles/kernel_sweep_r01.md, profiles/r02_kbench_k1_tiles.log).  Slot 1 in fp64,
  // strict arithmetic: 2 rows x 16 warps with the split barrier is the fastest tile at every fused depth
  // (T = 4: 0.892 vs 0.917 ms per pass at 512^3, T = 3: 0.700 vs 0.848, T = 2: 0.562 vs 0.648)
  if constexpr (K == 1 && sizeof(R) == 8 && T >= 2) return launch_r1_t<K, R, T, 2, 16, R1_SPLIT>(s);
  // fp32 at depth 4: the same tile with the split barrier, 1 040 against 1 013 GLUP/s (depths 2 and 3: no gain)
  if constexpr (K == 1 && sizeof(R) == 4 && T == 4) return launch_r1_t<K, R, T, 2, 16, R1_SPLIT>(s);
  // per-point coefficients at depth 2: the trapezoid skip (outer warps skip the last level and its coefficient loads) is
  // the fastest tile for every fp64 operator and for slot 5 in fp32 (profiles/r02_kbench_others.log)
  if constexpr (KTraits<K>::NCA > 0 && T == 2 && (sizeof(R) == 8 || K == 5)) return launch_r1_t<K, R, T, 2, 16, R1_TRAP>(s);
  // otherwise fp64 constant coefficients run best with 4 rows per thread and 8 warps (fewest shared-memory
  // exchanges per point), everything else with 2 rows and 16 warps
  if constexpr (KTraits<K>::NCA == 0 && sizeof(R) == 8) return launch_r1_t<K, R, T, 4, 8>(s);
  else return launch_r1_t<K, R, T, 2, 16>(s);
}

// single-step pass: the barrier-free marching kernel
template <int K, typename R, int PY, int NWY, bool FM = false>
static cudaError_t launch_march_t(const StreamLaunch &s) {
  const DevGrid &g = s.g;
  constexpr int WX = 32 * Vec<R>::N;
  MarchArgs<R> a;
  a.g = g;
  a.in = (const R *)s.in;
  a.out = (R *)s.out;
  a.coef = (const R *)s.coef;
  a.coef_stride = s.coef_stride;
  a.cc = make_cc<R>(s.cc);
  a.zb0 = s.zb0;
  a.ze0 = s.ze0;
  const int nz = s.ze0 - s.zb0;
  int zchunk = s.zchunk > 0 ? s.zchunk : 64;
  // every chunk re-reads two planes; keep chunks long but leave >= ~8 waves of CTAs
  constexpr int TY = PY * NWY;
  const int ntiles = ((g.nx + WX - 1) / WX) * ((g.ny + TY - 1) / TY);
  if (s.zchunk <= 0) {
    while (zchunk < nz && (long long)ntiles * ((nz + zchunk - 1) / zchunk) > 148LL * 8 * 8) zchunk *= 2;
    while (zchunk > 16 && (long long)ntiles * ((nz + zchunk - 1) / zchunk) < 148LL * 8 * 2) zchunk /= 2;
  }
  zchunk = std::min(zchunk, std::max(nz, 1));
  a.zchunk = zchunk;
  dim3 grid((g.nx + WX - 1) / WX, (g.ny + TY - 1) / TY, (nz + zchunk - 1) / zchunk);
  auto kfn = k_r1_march<K, R, PY, NWY, FM>;
  GIRIH_LAUNCH(kfn, grid, 32 * NWY, 0, s.stream, a);
  return cudaGetLastError();
}

template <int K, typename R>
static cudaError_t launch_r1_depth(int T, const StreamLaunch &s) {
  // a pass that pushes its boundary planes into the neighbours' halos needs the fused-sweep kernel (R1_PUSH) also
  // at depth 1: the marching kernel has no push stores
  const bool pushes = (s.push_up != nullptr) || (s.push_dn != nullptr);
  if (T == 1 && s.variant != 2 && !pushes) {
    if (s.contract) {
      if constexpr (K == 1 || K == 5) return launch_march_t<K, R, 4, 4, true>(s);
      else return launch_march_t<K, R, 2, 8, true>(s);
    }
    // tile option for the marching kernel: PY*100 + NWY
    if (s.tile == 108) return launch_march_t<K, R, 1, 8>(s);
    if (s.tile == 208) return launch_march_t<K, R, 2, 8>(s);
    if (s.tile == 404) return launch_march_t<K, R, 4, 4>(s);
    if (s.tile == 408) return launch_march_t<K, R, 4, 8>(s);
    if constexpr (K == 1 || K == 5) return launch_march_t<K, R, 4, 4>(s);
    else return launch_march_t<K, R, 2, 8>(s);
  }
  switch (T) {
    case 1: return launch_r1_tile<K, R, 1>(s);
    case 2: return launch_r1_tile<K, R, 2>(s);
    case 3: return launch_r1_tile<K, R, 3>(s);
    case 4:
      if constexpr (KTraits<K>::NCA == 0) return launch_r1_tile<K, R, 4>(s);
      else return cudaErrorInvalidValue;
    default: return cudaErrorInvalidValue;
  }
}

#define GIRIH_INST_R1(K, R, SUFFIX) \
  cudaError_t launch_r1_##SUFFIX(int T, const StreamLaunch &s) { return launch_r1_depth<K, R>(T, s); }

}  // namespace girih
