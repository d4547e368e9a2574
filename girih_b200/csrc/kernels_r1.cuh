// kernels_r1.cuh -- z-streamed, temporally fused sweep for the radius-1 star operators
// (table slots 1, 2, 3, 5).  T = 1 is the single-step stepper (ts 0/1), T > 1 the GPU form of
// GIRIH's wavefront-diamond sweep (ts 2; src/kernels/stencils_1wf.ic:33-82 and
// src/kernels/diamond_ts.c:427-488 describe the CPU schedule this replaces).
//
// Schedule (one CTA = NW warps):
//   * a lane owns VX = 16 B / sizeof(Real) consecutive x points, so every global access is a
//     coalesced 128-bit load/store; a warp spans WX = 32*VX points of one row
//   * a thread owns PY consecutive rows -> register tile VX x PY; the CTA tile is WX x (NW*PY)
//   * the CTA streams along z.  For every fused level l < T each thread keeps its own points of
//     the two newest planes of that level in registers (B = z-2, C = z-1 relative to the plane F
//     that level l produces in this iteration): the z neighbours never touch shared memory
//   * x neighbours inside a lane come from registers, across lanes from warp shuffles
//   * y neighbours inside a thread come from registers; only the first/last row of each warp's
//     strip goes through shared memory (double buffered, ONE __syncthreads per z iteration for
//     all T levels)
//   * level l+1 lags level l by one plane (the reference's `kt -= NHALO` skew,
//     stencils_1wf.ic:77), so T time steps cost one read and one write of the grid
//   * tiles overlap by HX >= T columns and T rows per side (halo depth = steps x radius);
//     whatever a tile computes outside its core is discarded.  Points that are not interior
//     points of the GLOBAL domain pass their value through every level, which keeps the
//     Dirichlet frame (and the +100.1 source planes, src/utils.c:679-696) exactly as the
//     reference, which simply never writes them (xb = r .. xe = nx + r).
#pragma once
#include "common.cuh"
#include "stencil_expr.cuh"

// kernels_r1x.cuh writes its loop body as nested lambdas that are too large for the inliner's own budget (a call would
// put the register planes on the stack).  NOT applied to k_r1 below: forcing its lambdas costs 4% (measured, round 2:
// 0.956 vs 0.917 ms per fp64 T = 4 pass; 254 instead of 248 registers and a different schedule)
#define GIRIH_LAMBDA_INLINE __attribute__((always_inline))

namespace girih {

template <typename R> struct R1Args {
  DevGrid g;
  const R *__restrict__ in;     // level L   (read)
  R *__restrict__ out;          // level L+T (written, interior of this slab only)
  const R *__restrict__ coef;   // per-point coefficient arrays (slots 2,3,5) or nullptr
  long long coef_stride;
  ConstCoef<R> cc;              // scalar coefficients (slot 1)
  int zb0, ze0;                 // output planes [zb0, ze0) of this launch (device z)
  int zchunk;                   // output planes per CTA
  int nch0;                     // z chunks (blockIdx.z) that belong to [zb0, ze0); the rest cover a second
  int zb1, ze1;                 // range [zb1, ze1): one launch can sweep the two outer parts of a slab
  // Halo push (R1_PUSH kernels): the output planes the z neighbours read as their halo in the next pass are
  // also stored straight into the neighbours' arrays over NVLink (peer memory), so a fused pass needs no
  // separate exchange.  push_up / push_dn point at the neighbour's output array, shifted so that indexing
  // them with MY plane number lands in its halo; planes >= push_up_from go up, planes < push_dn_below go down.
  R *push_up, *push_dn;
  int push_up_from, push_dn_below;
};

template <typename R, int T, int PY, int NW> struct R1Cfg {
  static constexpr int VX = Vec<R>::N;
  static constexpr int WX = 32 * VX;
  static constexpr int HX = ((T + VX - 1) / VX) * VX;   // x overlap, kept 16-byte aligned
  static constexpr int UX = WX - 2 * HX;                // core columns per tile
  static constexpr int H = NW * PY;
  static constexpr int UY = H - 2 * T;                  // core rows per tile
  static constexpr int PF = (T == 1) ? 2 : 1;           // planes prefetched ahead
  static constexpr int MINB = (T == 1 && NW <= 8) ? 2 : 1;   // resident CTAs per SM aimed for
  static constexpr size_t EDGE_BYTES = (size_t)T * 2 * NW * 2 * WX * sizeof(R);
  static constexpr size_t SMEM = EDGE_BYTES + 16;   // + the mbarrier of the split-barrier variant
  static_assert(UY > 0 && UX > 0, "tile too small for this fusion depth");
};

// neighbour accessor over registers: centre plane C, plane below B, plane above F
template <typename R> struct RegNb1 {
  R c, xm, xp, ym, yp, zm, zp;
  template <int DX, int DY, int DZ> __device__ __forceinline__ R at() const {
    if constexpr (DX == 0 && DY == 0 && DZ == 0) return c;
    else if constexpr (DX == -1) return xm;
    else if constexpr (DX == 1) return xp;
    else if constexpr (DY == -1) return ym;
    else if constexpr (DY == 1) return yp;
    else if constexpr (DZ == -1) return zm;
    else return zp;
  }
};

constexpr int R1_FM = 32;   // DBG flag bit: evaluate with the reference's gcc -mfma contraction
constexpr int R1_SPLIT = 64;   // DBG flag bit: split (arrive ... wait) CTA barrier instead of __syncthreads
constexpr int R1_REV = 128;    // DBG flag bit: decoupled levels (T > 1), see the comment in k_r1
constexpr int R1_PUSH = 256;   // DBG flag bit: boundary output planes are also stored into the z neighbours' halos
constexpr int R1_TRAP = 512;   // DBG flag bit: warps outside the rows a level is needed on skip that level (see k_r1)

// Split CTA barrier on an mbarrier in shared memory: a warp ARRIVES as soon as it has published its edge rows
// and read its neighbours' (before the arithmetic of the last fused level and the global stores) and WAITS at
// the top of the next iteration, so the warps of a CTA may drift apart by that much instead of meeting at one
// point per iteration.  arrive = release, wait = acquire (CTA scope).
#ifndef GIRIH_CUDA_EMU
__device__ __forceinline__ void sb_init(unsigned long long *bar, int count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void sb_arrive(unsigned long long *bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void sb_wait(unsigned long long *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n.reg .pred p;\nSB_WAIT:\nmbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n@p bra SB_DONE;\n"
      "bra SB_WAIT;\nSB_DONE:\n}\n" ::"r"(a), "r"(parity) : "memory");
}
#endif   // the test suite's CPU SIMT emulator (tests/cuda_emu) supplies its own versions

template <int PH> struct Phase { static constexpr int value = PH; };
template <int I> struct Level { static constexpr int value = I; };
template <bool B> struct FrameTag { static constexpr bool value = B; };
template <int I, int N, typename F> __device__ __forceinline__ void static_for(F &&f) {
  if constexpr (I < N) {
    f(Level<I>{});
    static_for<I + 1, N>(f);
  }
}
template <int N, typename F> __device__ __forceinline__ void static_for_rev(F &&f) {   // N-1, ..., 0
  if constexpr (N > 0) {
    f(Level<N - 1>{});
    static_for_rev<N - 1>(f);
  }
}

// halo push: one thread's rows of an output plane once more, into a neighbour's halo plane (peer memory).  A
// force-inlined function, not a lambda: a call the inliner declines would put the kernel on the ABI path (stack frame,
// 71 registers).
template <typename R, int PY, int VX>
__device__ __forceinline__ void push_rows(R *qp, const R (&O)[PY][VX], unsigned full_rows, unsigned part_rows,
                                          unsigned core_xy, int px, bool any_part) {
#pragma unroll
  for (int j = 0; j < PY; ++j)
    if ((full_rows >> j) & 1u) st128<R>(qp + (long long)j * px, O[j]);
  if (any_part) {   // rows cut by the domain edge (nx not a multiple of VX)
#pragma unroll
    for (int j = 0; j < PY; ++j) {
      if ((part_rows >> j) & 1u) {
#pragma unroll
        for (int e = 0; e < VX; ++e)
          if ((core_xy >> (j * VX + e)) & 1u) qp[(long long)j * px + e] = O[j][e];
      }
    }
  }
}

template <int K, typename R, int T, int PY, int NW, int DBG = 0>
__global__ void __launch_bounds__(32 * NW, R1Cfg<R, T, PY, NW>::MINB)
k_r1(const R1Args<R> a) {
  using Cfg = R1Cfg<R, T, PY, NW>;
  constexpr int VX = Cfg::VX, WX = Cfg::WX, HX = Cfg::HX, UX = Cfg::UX, H = Cfg::H, UY = Cfg::UY;
  constexpr int PF = Cfg::PF;
  constexpr int NCA = KTraits<K>::NCA;
  constexpr bool FM = (DBG & R1_FM) != 0;   // contracted (FMA) arithmetic, see stencil_expr.cuh
  constexpr bool SPLIT = (DBG & R1_SPLIT) != 0;
  // Decoupled levels (REV): level l+1 lags level l by TWO planes instead of one and the levels of an iteration
  // run from the deepest to the shallowest.  Level l+1 then consumes what level l produced in the PREVIOUS
  // iteration, so the T updates of one iteration do not depend on each other (the scheduler may interleave
  // them freely), the plane a level produces goes into the register plane its consumer has just released
  // (no extra registers), and the level-0 plane loaded at the end of one iteration is first touched at the
  // end of the next one.  Price: T-1 more pipeline-fill iterations per z chunk.
  constexpr bool REV = (DBG & R1_REV) != 0 && T > 1;
  constexpr int LAG = REV ? 2 : 1;
  static_assert(!(REV && SPLIT), "not combined");
  // Trapezoid skip (TRAP): the plane the last level produces is only needed on the core rows [T, H-T) of the tile,
  // so a warp whose PY rows all lie outside the core skips the last level and the stores altogether (warp-uniform
  // branch around arithmetic + stores: no shuffles, no shared-memory reads, nothing live across it).  Only the
  // LAST level is skipped: branches around the inner levels split the fused basic block and spill.  With PY = 4,
  // T = 4 the first and last warp skip 1 of 4 levels (1/16 of the tile's updates), with PY = 2, T = 4 two warps on
  // each side do (also 1/16).
  constexpr bool TRAP = (DBG & R1_TRAP) != 0 && T > 1 && PY <= T;
  static_assert(!(TRAP && (REV || SPLIT)), "not combined");
  constexpr unsigned ALL = (PY * VX >= 32) ? 0xffffffffu : ((1u << (PY * VX)) - 1u);
  static_assert(KTraits<K>::R == 1 && KTraits<K>::TO == 1, "radius-1, first-order-in-time only");
  static_assert(PY * VX <= 32, "point masks are 32 bits");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // edge[l][parity][warp][0 = first row, 1 = last row][WX]
  R *edge = reinterpret_cast<R *>(smem_raw);
  unsigned long long *const sbar = reinterpret_cast<unsigned long long *>(smem_raw + Cfg::EDGE_BYTES);
  if constexpr (SPLIT) {
    if (threadIdx.x == 0) sb_init(sbar, 32 * NW);
    __syncthreads();
  }
  auto edge_ptr = [&](int l, int par, int w, int which) -> R * {
    return edge + ((((size_t)l * 2 + par) * NW + w) * 2 + which) * WX;
  };

  const DevGrid &g = a.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = g.X0 - HX + (int)blockIdx.x * UX + lane * VX;   // my first column (device x)
  const int y0 = g.Y0 - T + (int)blockIdx.y * UY + warp * PY;   // my first row
  const int cz = (int)blockIdx.z;
  const bool second = cz >= a.nch0;
  const int zb = second ? a.zb1 + (cz - a.nch0) * a.zchunk : a.zb0 + cz * a.zchunk;
  const int ze = min(zb + a.zchunk, second ? a.ze1 : a.ze0);

  // per-point facts that do not change along z
  const bool x_alloc = (x >= 0) && (x + VX <= g.px);
  unsigned row_alloc = 0;     // bit j: row j of my strip may be loaded
  unsigned interior_xy = 0;   // bit j*VX+e: point is an interior point in x and y
  unsigned core_xy = 0;       // ... and lies in this tile's core -> this tile stores it
#pragma unroll
  for (int j = 0; j < PY; ++j) {
    const int y = y0 + j;
    // rows/columns past the domain (partial last tiles) are not read: nothing valid depends on them
    if (x_alloc && (y >= 0) && (y < g.ny_dev) && (y < g.Y0 + g.ny + g.r) && (x < g.X0 + g.nx + g.r))
      row_alloc |= 1u << j;
    const bool yin = (y >= g.Y0) && (y < g.Y0 + g.ny);
    const int ly = warp * PY + j;
    const bool ycore = (ly >= T) && (ly < H - T);
#pragma unroll
    for (int e = 0; e < VX; ++e) {
      const bool xin = (x + e >= g.X0) && (x + e < g.X0 + g.nx);
      const int lx = lane * VX + e;
      const bool xcore = (lx >= HX) && (lx < WX - HX);
      if (xin && yin) interior_xy |= 1u << (j * VX + e);
      if (xin && yin && xcore && ycore) core_xy |= 1u << (j * VX + e);
    }
  }
  // warps that lie completely inside the domain skip the per-point pass-through fix-up
  const bool warp_masked = __any_sync(0xffffffffu, interior_xy != ALL);
  const bool outer_warp = ((warp + 1) * PY <= T) || (warp * PY >= H - T);   // no core row (TRAP)
  unsigned full_rows = 0, part_rows = 0;   // rows stored whole / rows stored point by point
#pragma unroll
  for (int j = 0; j < PY; ++j) {
    const unsigned m = (core_xy >> (j * VX)) & ((1u << VX) - 1u);
    if (m == (1u << VX) - 1u) full_rows |= 1u << j;
    else if (m != 0u) part_rows |= 1u << j;
  }
  const bool any_part = __any_sync(0xffffffffu, part_rows != 0u);
  const long long row0 = (long long)y0 * g.px + x;   // offset of my first point inside a plane
  const R *const coef_t = a.coef + row0;   // per-thread base; the (z, row) offset below is warp-uniform
  const int nit = (ze - zb) + (REV ? 3 * T - 1 : 2 * T);

  // The loop body exists twice: FRAME = false for iterations in which this warp cannot see a
  // non-interior point (no pass-through code at all), FRAME = true for warps on the x/y frame and for
  // the few iterations of the first/last z chunk whose planes touch the z frame.  Both versions execute
  // exactly one __syncthreads per iteration, so the warps of a CTA may take different versions.
  // (The warps then reach the barrier from different code addresses.  sm_70+ counts arrivals per barrier, not per
  // instruction, and every thread of the CTA executes exactly one bar.sync 0 per iteration on either path; the test
  // suite's emulator counts the same way, and `compute-sanitizer --tool synccheck` on the GPU suite's fused cases is
  // clean -- profiles/r02_synccheck.log.)
  {
  // S[l][.] : three rotating register planes of level l.  In phase PH (= iteration mod 3)
  //   S[l][PH] = plane zc-1 ("B"), S[l][(PH+1)%3] = plane zc ("C"), S[l][(PH+2)%3] = plane zc+1 ("F")
  // where zc is the plane level l+1 is produced at in this iteration.  Level l+1 writes its new plane
  // straight into S[l+1][(PH+2)%3]; the roles rotate with the unrolled phase, so no register moves.
  R S[T][3][PY][VX];
#pragma unroll
  for (int l = 0; l < T; ++l)
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int e = 0; e < VX; ++e) S[l][q][j][e] = (R)0;

  // T == 1: PF planes are prefetched into a small queue.  T > 1: the next level-0 plane is loaded
  // straight into the register plane that becomes "F" in the next phase (this phase's "B" of level 0,
  // dead once level 1 has been produced), so the loads fly underneath levels 2..T.
  R nxt[T == 1 ? PF : 1][PY][VX];
#pragma unroll
  for (int q = 0; q < (T == 1 ? PF : 1); ++q)
#pragma unroll
    for (int j = 0; j < PY; ++j)
#pragma unroll
      for (int e = 0; e < VX; ++e) nxt[q][j][e] = (R)0;

  // branch-free (predicated loads): the fused levels stay one basic block, so the scheduler can run
  // one level's arithmetic under another level's shuffle / shared-memory latency
  auto load_plane = [&](int z, R (&dst)[PY][VX], bool want = true) {
    const unsigned m = (want && (z >= 0) && (z < g.nz_dev)) ? row_alloc : 0u;
    const R *p = a.in + (long long)z * g.pxy + row0;
#pragma unroll
    for (int j = 0; j < PY; ++j)
      if ((m >> j) & 1u) ld128<R>(p + (long long)j * g.px, dst[j]);
  };
  if constexpr (T == 1) {
#pragma unroll
    for (int q = 0; q < PF; ++q) load_plane(zb - T + q, nxt[q]);
  } else {
    load_plane(zb - T, S[0][2]);   // "F" of phase 0
  }


  auto body = [&](auto phase_tag, auto frame_tag, const int it) {
    constexpr int PH = decltype(phase_tag)::value;
    constexpr bool FRAME = decltype(frame_tag)::value;
    constexpr int iB = PH, iC = (PH + 1) % 3, iF = (PH + 2) % 3;
    const int zin = zb - T + it;
    const int cur = it & 1;
    // split barrier: everything published in iteration it-1 is visible, and everyone has finished reading the
    // buffers this iteration overwrites
    if constexpr (SPLIT) { if (it > 0) sb_wait(sbar, (unsigned)((it - 1) & 1)); }

    if constexpr (T == 1) {
      // level 0: take the prefetched plane, keep the prefetch queue full
#pragma unroll
      for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int e = 0; e < VX; ++e) S[0][iF][j][e] = nxt[0][j][e];
#pragma unroll
      for (int q = 0; q + 1 < PF; ++q)
#pragma unroll
        for (int j = 0; j < PY; ++j)
#pragma unroll
          for (int e = 0; e < VX; ++e) nxt[q][j][e] = nxt[q + 1][j][e];
      load_plane(zin + PF, nxt[PF - 1], it + PF < nit);
    }

    // publish the first/last row of the level-0 plane for next iteration's level-1 update
    if constexpr (!(DBG & 2) && !REV) {
      st128<R>(edge_ptr(0, cur, warp, 0) + lane * VX, S[0][iF][0]);
      st128<R>(edge_ptr(0, cur, warp, 1) + lane * VX, S[0][iF][PY - 1]);
    }

    R Ofin[PY][VX];
    auto level = [&](auto level_tag) {
      constexpr int l = decltype(level_tag)::value;
      // level l+1 at plane zc from level l planes zc-1 (B), zc (C), zc+1 (F)
      const int zc = zin - LAG * l - 1;
      R (&Bp)[PY][VX] = S[l][iB];
      R (&Cp)[PY][VX] = S[l][iC];
      R (&Fp)[PY][VX] = S[l][iF];
      // REV: every level publishes the edge rows of its own newest plane (the centre plane of the next
      // iteration); level 0 does so here too, i.e. at the end of the iteration, after its load has landed
      if constexpr (REV) {
        st128<R>(edge_ptr(l, cur, warp, 0) + lane * VX, Fp[0]);
        st128<R>(edge_ptr(l, cur, warp, 1) + lane * VX, Fp[PY - 1]);
      }

      // rows just outside my strip, owned by the neighbouring warps (written last iteration)
      R up[VX], dn[VX];
      {
        const int wu = (warp > 0) ? warp - 1 : 0, wd = (warp < NW - 1) ? warp + 1 : NW - 1;
        if constexpr (DBG & 2) {   // experiment: no shared-memory exchange (results invalid)
#pragma unroll
          for (int e = 0; e < VX; ++e) { up[e] = Cp[0][e]; dn[e] = Cp[PY - 1][e]; }
        } else {
          ld128s<R>(edge_ptr(l, cur ^ 1, wu, 1) + lane * VX, up);
          ld128s<R>(edge_ptr(l, cur ^ 1, wd, 0) + lane * VX, dn);
        }
        // last shared-memory access of this iteration (the edges of levels 1..T-1 were published by the
        // earlier stages, level T is not exchanged): arrive now, run this level and the stores underneath
        if constexpr (SPLIT && l == T - 1) sb_arrive(sbar);
      }

      auto stage = [&](R (&O)[PY][VX]) {
#pragma unroll
      for (int j = 0; j < PY; ++j) {
        // per-point coefficients of plane zc, row j (slots 2, 3, 5)
        R cf[NCA > 0 ? NCA : 1][VX];
        if constexpr (NCA > 0) {
          const bool ok = (zc >= 0) && (zc < g.nz_dev) && ((row_alloc >> j) & 1u);
          const R *cp = coef_t + ((long long)zc * g.pxy + (long long)j * g.px);
#pragma unroll
          for (int m = 0; m < NCA; ++m) {
#pragma unroll
            for (int e = 0; e < VX; ++e) cf[m][e] = (R)0;
            if (ok) ld128<R>(cp + (long long)m * a.coef_stride, cf[m]);
          }
        }
        const R left = (DBG & 1) ? Cp[j][0] : __shfl_up_sync(0xffffffffu, Cp[j][VX - 1], 1);
        const R right = (DBG & 1) ? Cp[j][VX - 1] : __shfl_down_sync(0xffffffffu, Cp[j][0], 1);
#pragma unroll
        for (int e = 0; e < VX; ++e) {
          RegNb1<R> n;
          n.c = Cp[j][e];
          n.xm = (e > 0) ? Cp[j][e > 0 ? e - 1 : 0] : left;
          n.xp = (e < VX - 1) ? Cp[j][e < VX - 1 ? e + 1 : 0] : right;
          n.ym = (j > 0) ? Cp[j > 0 ? j - 1 : 0][e] : up[e];
          n.yp = (j < PY - 1) ? Cp[j < PY - 1 ? j + 1 : 0][e] : dn[e];
          n.zm = Bp[j][e];
          n.zp = Fp[j][e];
          if constexpr (NCA > 0) {
            RegCoef<R, NCA> rc;
#pragma unroll
            for (int m = 0; m < NCA; ++m) rc.v[m] = cf[m][e];
            O[j][e] = StencilExpr<K>::template eval<R, FM>(n, rc, (R)0, (R)0);
          } else {
            O[j][e] = StencilExpr<K>::template eval<R, FM>(n, a.cc, (R)0, (R)0);
          }
        }
      }
        // pass-through of everything that is not an interior point of the global domain: a frame
        // plane clears the whole mask (CTA-uniform), the x/y frame clears single bits
        if constexpr (FRAME) {
          const unsigned upd = ((zc >= g.zlo) && (zc < g.zhi)) ? interior_xy : 0u;
#pragma unroll
          for (int j = 0; j < PY; ++j)
#pragma unroll
            for (int e = 0; e < VX; ++e) O[j][e] = ((upd >> (j * VX + e)) & 1u) ? O[j][e] : Cp[j][e];
        }
        if constexpr (l + 1 < T && !(DBG & 2) && !REV) {
          st128<R>(edge_ptr(l + 1, cur, warp, 0) + lane * VX, O[0]);
          st128<R>(edge_ptr(l + 1, cur, warp, 1) + lane * VX, O[PY - 1]);
        }
      };
      // forward order: the new plane is this iteration's "F" of level l+1.  REV: level l+1 has already run in
      // this iteration and released its "B" plane, which is next iteration's "F".
      if constexpr (l + 1 < T) stage(S[l + 1][REV ? iB : iF]);
      else stage(Ofin);
      if constexpr (T > 1 && l == 0) {
        // next iteration's level-0 "F"; REV streams T-1 iterations longer than there are planes to read
        load_plane(zin + 1, S[0][iB], (it + 1 < nit) && (!REV || zin + 1 < ze + T));
      }
    };
    auto store_out = [&]() {
    // Ofin is level T at plane zin - LAG*(T-1) - 1.  Rows whose VX points all lie in the core go out as predicated
    // 128-bit stores (no branches); rows cut by the domain edge (nx not a multiple of VX) are rare and
    // take a warp-uniform slow path.
    const int zo = zin - LAG * (T - 1) - 1;
    const bool zst = (zo >= zb) && (zo < ze);
    R *q = a.out + (long long)zo * g.pxy + row0;
    const unsigned fm = zst ? full_rows : 0u;
#pragma unroll
    for (int j = 0; j < PY; ++j)
      if ((fm >> j) & 1u) st128<R>(q + (long long)j * g.px, Ofin[j]);
    if (any_part && zst) {
#pragma unroll
      for (int j = 0; j < PY; ++j) {
        if ((part_rows >> j) & 1u) {
          const unsigned m = (core_xy >> (j * VX)) & ((1u << VX) - 1u);
#pragma unroll
          for (int e = 0; e < VX; ++e)
            if ((m >> e) & 1u) q[(long long)j * g.px + e] = Ofin[j][e];
        }
      }
    }
    // halo push: the same rows once more, into the upper / lower neighbour's halo planes (peer memory).  Only the
    // few iterations that produce boundary planes enter (warp-uniform test on the plane number).
    if constexpr ((DBG & R1_PUSH) != 0) {
      if (zst && zo >= a.push_up_from)
        push_rows<R, PY, VX>(a.push_up + (long long)zo * g.pxy + row0, Ofin, full_rows, part_rows, core_xy, g.px, any_part);
      if (zst && zo < a.push_dn_below)
        push_rows<R, PY, VX>(a.push_dn + (long long)zo * g.pxy + row0, Ofin, full_rows, part_rows, core_xy, g.px, any_part);
    }
    };
    if constexpr (REV) {
      static_for_rev<T>(level);
      store_out();
    } else if constexpr (TRAP) {
      static_for<0, T - 1>(level);
      if (!outer_warp) {
        level(Level<T - 1>{});
        store_out();
      }
    } else {
      static_for<0, T>(level);
      store_out();
    }
    if constexpr (!(DBG & 2) && !SPLIT) __syncthreads();
  };

  // planes zin-T .. zin-1 are produced in iteration `it`; the version is chosen per iteration
  auto step = [&](auto phase_tag, const int it) {
    const int zin = zb - T + it;
    if constexpr (DBG & 8) body(phase_tag, FrameTag<false>{}, it);       // experiment (results invalid)
    else if constexpr (DBG & 16) body(phase_tag, FrameTag<true>{}, it);   // experiment
    else if (warp_masked || (zin - LAG * (T - 1) - 1 < g.zlo) || (zin > g.zhi)) body(phase_tag, FrameTag<true>{}, it);
    else body(phase_tag, FrameTag<false>{}, it);
  };
  if constexpr ((DBG & R1_PUSH) != 0) {
    // one copy of each phase instead of loop + remainder copies: with the push stores the fully unrolled form
    // outgrows the inliner's budget in fp32 and the bodies would become real calls (stack frame, 71 registers)
    for (int it = 0; it < nit;) {
      step(Phase<0>{}, it++);
      if (it >= nit) break;
      step(Phase<1>{}, it++);
      if (it >= nit) break;
      step(Phase<2>{}, it++);
    }
  } else {
  int it = 0;
  for (; it + 3 <= nit; it += 3) {
    step(Phase<0>{}, it);
    step(Phase<1>{}, it + 1);
    step(Phase<2>{}, it + 2);
  }
  if (it < nit) { step(Phase<0>{}, it); ++it; }
  if (it < nit) { step(Phase<1>{}, it); }
  }
  }
}

}  // namespace girih
