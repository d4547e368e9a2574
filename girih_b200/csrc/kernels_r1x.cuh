// kernels_r1x.cuh -- temporally fused sweep of the radius-1 star operators on NON-OVERLAPPING tiles.
//
// k_r1 (kernels_r1.cuh) gives every CTA a tile that overlaps its neighbours by T*r points per side and
// recomputes the overlap: at T = 4 a 64 x 32 tile keeps 65.6% of what it computes (56.5% at 512^2 with
// the partial tiles).  This kernel is the GPU form of what GIRIH's diamond tiling is for
// (src/kernels/diamond_ts.c:565-595, intra_diamond_get_info_std: yb, ye, b_inc, e_inc; the y extent of a
// tile changes by r per level, stencils_1wf.ic:65-71) -- no point is computed twice -- expressed for a
// machine whose "thread groups" are 148 co-resident CTAs with a 126 MB L2 between them:
//
//   * the x-y plane is cut into EXACT tiles of WX x H points (no overlap), one CTA per tile and z chunk,
//     all CTAs of the launch co-resident (cooperative launch); a CTA streams along z with the same
//     register pipeline as k_r1 (three rotating register planes per fused level, level l+1 one plane
//     behind level l)
//   * what a tile is missing at its rim -- for every level 1 <= l < T the column left / right of the tile and
//     the row above / below it, plane by plane -- is handed over by the neighbouring CTA as soon as it has
//     produced it: the producer stores the values straight into the consumer's inbound slots in global
//     memory (they live in L2), every 8-byte word carrying its own sequence tag (payload, tag) like
//     NCCL's LL protocol, so there is no fence, no flag round trip and no barrier between CTAs; the
//     consumer polls a slot until the tag of the iteration it needs shows up.  Level 0 is the input array:
//     its rim is read from global memory by every tile, like the frame cells of the tiles on the Dirichlet frame
//   * a value is published in iteration g (stage l-1) and consumed in iteration g+1 (stage l): one whole
//     iteration of slack against a measured hand-over latency of 700-1050 cycles; slots are a ring of three
//     generations, which the dependence chain makes sufficient
//   * every poll is private to the consuming warp and issued in the stage before the one that needs the value
//     (T-1 stages after the store), one 16-byte slot per lane in flight, checked after the stage's arithmetic
//   * every store / load of the exchange in the hot block is a predicated instruction, not a branch (the T fused
//     levels stay one basic block); rim columns are parked in shared memory by lanes 0 / 31 and leave with ONE
//     store per warp and stage; the rows above / below leave with the first / last warp
// Status: bit-exact (GPU + emulator), NOT a performance path -- 0.60 ms per pass without the exchange (889 GLUP/s
// at 512^3) but 2.3-3.4 ms with it; the cost of each part is in profiles/r02_exact_tiles_analysis.md.
//
// fp64 only (one value + tags = one 16-byte slot).
#pragma once
#include "common.cuh"
#include "kernels_r1.cuh"
#include "stencil_expr.cuh"

namespace girih {

template <typename R> struct R1xArgs {
  DevGrid g;
  const R *__restrict__ in;
  R *__restrict__ out;
  const R *__restrict__ coef;
  long long coef_stride;
  ConstCoef<R> cc;
  int zb0, ze0;          // output planes [zb0, ze0) (device z)
  int zchunk;            // output planes per CTA (blockIdx.z)
  unsigned char *xbuf;   // inbound edge slots, R1xCfg::TILE_BYTES per CTA (linear block index)
  unsigned seq0;         // iteration g of this launch tags its slots with seq0 + g + 1
  int *err;              // set to 1 when a poll gave up (a neighbour tile never delivered): the host reports it
#ifdef GIRIH_R1X_TRACE
  long long *trace;      // measurement build only: [cta][warp 0 / warp 3][iteration 96..111][stage][before / after the wait]
#endif
};

template <typename R, int T, int PY, int NW> struct R1xCfg {
  static_assert(sizeof(R) == 8, "fp64 only");
  static constexpr int VX = Vec<R>::N;
  static constexpr int WX = 32 * VX;
  static constexpr int H = NW * PY;
  static constexpr int NXS = 2 * PY;                       // x slots of one warp and level (both sides)
  static constexpr int LEVEL_SLOTS = 2 * H + 2 * WX;       // [x-: H][x+: H][y-: WX][y+: WX]
  static constexpr int RING_SLOTS = T * LEVEL_SLOTS;
  static constexpr int RING = 3;
  static constexpr size_t TILE_BYTES = (size_t)RING * RING_SLOTS * 16;
  static constexpr size_t EDGE_BYTES = (size_t)T * 2 * (NW + 2) * 2 * WX * sizeof(R);
  static constexpr size_t XS_BYTES = (size_t)NW * T * 2 * PY * sizeof(R);
  static constexpr size_t XO_BYTES = (size_t)NW * NXS * sizeof(R);
  static constexpr size_t SMEM = EDGE_BYTES + XS_BYTES + XO_BYTES;
  static_assert(T >= 2 && NW >= 2 && NXS <= 32, "one x slot per lane and stage; first and last warp are distinct");
};

// ---- LL slots: {payload lo, tag, payload hi, tag}; each 8-byte half is written / read atomically ----------
constexpr unsigned R1X_SPIN_LIMIT = 1u << 22;   // polls (~0.3 us each) before a lane gives up on a slot
#ifndef GIRIH_CUDA_EMU
struct LLWord { unsigned lo, t0, hi, t1; };
// All of these are PREDICATED single instructions, not branches: the T fused levels of an iteration must stay one
// basic block (the scheduler overlaps one level's shuffles and loads with another level's arithmetic), and every
// `if` around a store or a load would cut it.
template <int OFF> __device__ __forceinline__ void ll_store_if(void *p, double v, unsigned tag) {   // no-op when p == nullptr
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("{\n.reg .pred q;\nsetp.ne.u64 q, %0, 0;\n@q st.relaxed.gpu.global.v4.u32 [%0+%5], {%1, %2, %3, %4};\n}"
               ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag), "n"(OFF));
}
__device__ __forceinline__ LLWord ll_load(const void *p) {
  LLWord w;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.lo), "=r"(w.t0), "=r"(w.hi), "=r"(w.t1) : "l"(p));
  return w;
}
__device__ __forceinline__ void ll_load_if(bool pred, const void *p, LLWord &w) {   // w keeps its value when !pred
  asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %5, 0;\n@q ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];\n}"
               : "+r"(w.lo), "+r"(w.t0), "+r"(w.hi), "+r"(w.t1) : "l"(p), "r"((unsigned)pred));
}
__device__ __forceinline__ void frame_load_if(bool pred, const double *p, double &v) {   // read-only path; v kept when !pred
  asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.global.nc.f64 %0, [%1];\n@q prefetch.global.L2 [%1+%3];\n}"
               : "+d"(v) : "l"(p), "r"((unsigned)pred), "n"(0));
}
__device__ __forceinline__ void prefetch_l2_if(bool pred, const void *p) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %1, 0;\n@q prefetch.global.L2 [%0];\n}" ::"l"(p), "r"((unsigned)pred));
}
__device__ __forceinline__ void lds_if(bool pred, const double *p, double &v) {   // shared memory; v kept when !pred
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.shared.f64 %0, [%1];\n}" : "+d"(v) : "r"(a), "r"((unsigned)pred));
}
__device__ __forceinline__ void sts_if(bool pred, double *p, double v) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.shared.f64 [%0], %1;\n}" ::"r"(a), "d"(v), "r"((unsigned)pred) : "memory");
}
__device__ __forceinline__ double ll_value(const LLWord &w) { return __hiloint2double((int)w.hi, (int)w.lo); }
__device__ __forceinline__ void spin_pause() { __nanosleep(40); }
#endif   // the test suite's CPU SIMT emulator supplies its own versions (tests/cuda_emu)

// XDBG (timing experiments only, results INVALID; built with -DGIRIH_PERF_EXPERIMENTS): 1 = no LL publish,
// 2 = no polls, 4 = rim columns not taken in (plain shuffles)
template <int K, typename R, int T, int PY, int NW, bool FM = false, int XDBG = 0>
__global__ void __launch_bounds__(32 * NW, 1)
k_r1x(const R1xArgs<R> a) {
  using Cfg = R1xCfg<R, T, PY, NW>;
  constexpr int VX = Cfg::VX, WX = Cfg::WX, H = Cfg::H, NXS = Cfg::NXS;
  constexpr int LEVEL_SLOTS = Cfg::LEVEL_SLOTS, RING_SLOTS = Cfg::RING_SLOTS;
  constexpr int NCA = KTraits<K>::NCA;
  constexpr unsigned ALL = (PY * VX >= 32) ? 0xffffffffu : ((1u << (PY * VX)) - 1u);
  static_assert(KTraits<K>::R == 1 && KTraits<K>::TO == 1, "radius-1, first-order-in-time only");
  static_assert(PY * VX <= 32, "point masks are 32 bits");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // edge[l][parity][warp][0 = first row, 1 = last row][WX]; "warps" NW and NW + 1 are the rows handed over by the
  // tiles above (y-) and below (y+)
  R *const edge = reinterpret_cast<R *>(smem_raw);
  auto edge_ptr = [&](int l, int par, int w, int which) GIRIH_LAMBDA_INLINE -> R * {
    return edge + ((((size_t)l * 2 + par) * (NW + 2) + w) * 2 + which) * WX;
  };
  const DevGrid &g = a.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // xs[level][side][row]: the columns left (side 0) / right (side 1) of this warp's rows, handed over by the x neighbours
  R *const xs = reinterpret_cast<R *>(smem_raw + Cfg::EDGE_BYTES) + (size_t)warp * (T * 2 * PY);

  const int bx = (int)blockIdx.x, by = (int)blockIdx.y, cz = (int)blockIdx.z;
  const int ni = (int)gridDim.x, nj = (int)gridDim.y;
  const int xt0 = g.X0 + bx * WX, yt0 = g.Y0 + by * H;   // first column / row of the tile
  const int x = xt0 + lane * VX;                          // my first column (device x)
  const int y0 = yt0 + warp * PY;                         // my first row
  const int zb = a.zb0 + cz * a.zchunk;
  const int ze = min(zb + a.zchunk, a.ze0);
  const bool has_xm = bx > 0, has_xp = bx + 1 < ni, has_ym = by > 0, has_yp = by + 1 < nj;
  const long long cta = ((long long)cz * nj + by) * ni + bx;
  unsigned char *const inb = a.xbuf + cta * (long long)Cfg::TILE_BYTES;   // my inbound slots

  // per-point facts that do not change along z
  const bool x_alloc = (x >= 0) && (x + VX <= g.px);
  unsigned row_alloc = 0, interior_xy = 0;
#pragma unroll
  for (int j = 0; j < PY; ++j) {
    const int y = y0 + j;
    if (x_alloc && (y >= 0) && (y < g.ny_dev) && (y < g.Y0 + g.ny + g.r) && (x < g.X0 + g.nx + g.r)) row_alloc |= 1u << j;
    const bool yin = (y >= g.Y0) && (y < g.Y0 + g.ny);
#pragma unroll
    for (int e = 0; e < VX; ++e) {
      const bool xin = (x + e >= g.X0) && (x + e < g.X0 + g.nx);
      if (xin && yin) interior_xy |= 1u << (j * VX + e);
    }
  }
  const bool warp_masked = __any_sync(0xffffffffu, interior_xy != ALL);
  unsigned full_rows = 0, part_rows = 0;   // rows stored whole / point by point (every interior point is stored)
#pragma unroll
  for (int j = 0; j < PY; ++j) {
    const unsigned m = (interior_xy >> (j * VX)) & ((1u << VX) - 1u);
    if (m == (1u << VX) - 1u) full_rows |= 1u << j;
    else if (m != 0u) part_rows |= 1u << j;
  }
  const bool any_part = __any_sync(0xffffffffu, part_rows != 0u);
  const long long row0 = (long long)y0 * g.px + x;
  const R *const coef_t = a.coef + row0;
  const int nit = (ze - zb) + 2 * T;

  // ---- publishing: where my rim values go -------------------------------------------------------------
  // warp 0 hands its first row to the tile above (its y+ slots), warp NW-1 its last row to the tile below (y- slots)
  unsigned char *ypub = nullptr;
  if (warp == 0 && has_ym) ypub = inb - (long long)ni * (long long)Cfg::TILE_BYTES + (size_t)(2 * H + WX + lane * VX) * 16;
  if (warp == NW - 1 && has_yp) ypub = inb + (long long)ni * (long long)Cfg::TILE_BYTES + (size_t)(2 * H + lane * VX) * 16;
  const bool edge_warp = (warp == 0) || (warp == NW - 1);
  const bool lane_first = lane == 0, lane_last = lane == 31;

  // ---- polling: the slots each lane collects ----------------------------------------------------------------
  // A rim value of level l is published in iteration g (when the stage that produces level l ends; level 0: at the top)
  // and consumed by stage l of iteration g + 1.  Every value is collected as LATE as possible, in the stage before the
  // one that consumes it: stage s of iteration `it` polls level s + 1 of iteration it - 1 (s < T - 1), stage T - 1 polls
  // level 0 of iteration `it`.  That leaves T - 1 stages (~1700 cycles at T = 4) between a store and the first look at
  // its slot, against a measured hand-over latency of 700 (same die) .. 1050 (other die) cycles
  // (tools/micro/ll_pingpong.cu), so a poll normally succeeds at once.  Everything is private to the consuming warp:
  //   x: lanes < NXS of every warp collect the warp's 2 * PY column values (one slot each) into xs[]
  //   y: the first / last warp collects the row above / below the tile, 2 slots per lane, into the edge rows
  // kinds: 0 none, 1 slot written by a neighbour tile, 2 Dirichlet frame cell read from the input array
  // Level 0 is the input array itself: its rim needs no hand-over, every tile reads it from global memory like a frame
  // cell (f0X / f0Y: is the cell allocated; fX / fY: its in-plane offset).  Only levels 1..T-1 travel through slots.
  int kX = 0, kY = 0;
  unsigned sX = 0, sY = 0;   // slot byte offset inside one (ring, level 0) block -- or in-plane element offset (kind 2)
  unsigned fX = 0, fY = 0;
  bool f0X = false, f0Y = false;
  int dX = 0, dY = 0;        // destination: xs index (level 0) / edge-row element offset (level 0, parity 0)
  unsigned char *xpub8 = nullptr;   // lanes < NXS: where rim value `lane` of this warp goes (null: no tile on that side)
  if (lane < NXS) {
    const int side = lane / PY, j = lane % PY;
    const bool has = side == 0 ? has_xm : has_xp;
    dX = side * PY + j;
    const int xc = side == 0 ? xt0 - 1 : xt0 + WX, yr = y0 + j;
    f0X = xc >= 0 && xc < g.px && yr >= 0 && yr < g.ny_dev;
    fX = (unsigned)(yr * g.px + xc);
    if (has) { kX = 1; sX = (unsigned)((side * H + warp * PY + j) * 16); }
    else { kX = f0X ? 2 : 0; sX = fX; }
    // my first column goes to the left tile's x+ slots, my last column to the right tile's x- slots
    if (has) xpub8 = (side == 0 ? inb - (long long)Cfg::TILE_BYTES + (size_t)H * 16 : inb + (long long)Cfg::TILE_BYTES) + (size_t)(warp * PY + j) * 16;
  }
  if (edge_warp) {
    const int side = warp == 0 ? 0 : 1;
    const bool has = side == 0 ? has_ym : has_yp;
    dY = (int)(edge_ptr(0, 0, NW + side, side == 0 ? 1 : 0) - edge) + lane * VX;
    const int yr = side == 0 ? yt0 - 1 : yt0 + H;
    f0Y = yr >= 0 && yr < g.ny_dev && x_alloc;
    fY = (unsigned)(yr * g.px + x);
    if (has) { kY = 1; sY = (unsigned)((2 * H + side * WX + lane * VX) * 16); }
    else { kY = f0Y ? 2 : 0; sY = fY; }
  }
  // rim values of the level a stage has just produced wait here (xo[side][row], written by lanes 0 / 31) until the warp's
  // first NXS lanes send them off with ONE store at the start of the next stage
  R *const xo = reinterpret_cast<R *>(smem_raw + Cfg::EDGE_BYTES + Cfg::XS_BYTES) + (size_t)warp * NXS;
  constexpr int EDGE_PAR = (NW + 2) * 2 * WX;        // elements between the two parities of one level's edge rows
  constexpr int EDGE_LVL = 2 * EDGE_PAR;             // ... between levels

  {
  R S[T][3][PY][VX];
#pragma unroll
  for (int l = 0; l < T; ++l)
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int e = 0; e < VX; ++e) S[l][q][j][e] = (R)0;

  auto load_plane = [&](int z, R (&dst)[PY][VX], bool want = true) {
    const unsigned m = (want && (z >= 0) && (z < g.nz_dev)) ? row_alloc : 0u;
    const R *p = a.in + (long long)z * g.pxy + row0;
#pragma unroll
    for (int j = 0; j < PY; ++j)
      if ((m >> j) & 1u) ld128<R>(p + (long long)j * g.px, dst[j]);
  };
  load_plane(zb - T, S[0][2]);   // "F" of phase 0

  // one slot: issue() starts the load (predicated instructions, no branch), finish() checks the tag and only then, in
  // the rare case that the value has not arrived, leaves the straight line to wait for it.  A frame cell (kind 2) is
  // read from plane zp of the input array and wrapped so that finish() accepts it at once.
  auto issue = [&](int kind, const unsigned char *slot, unsigned off, int zp, unsigned tag) GIRIH_LAMBDA_INLINE -> LLWord {
    LLWord w;
    w.lo = w.hi = 0u; w.t0 = w.t1 = tag;
    ll_load_if(kind == 1, slot, w);
    const bool fr = (kind == 2) && (zp >= 0) && (zp < g.nz_dev);
    const R *q = a.in + (long long)zp * g.pxy + off;
    R fv = (R)0;
    frame_load_if(fr, q, fv);
    prefetch_l2_if(fr && (zp + 2 < g.nz_dev), q + 2 * g.pxy);   // the same cell two planes on: wanted two iterations from now
    if (kind == 2) { w.lo = (unsigned)__double2loint(fv); w.hi = (unsigned)__double2hiint(fv); }
    return w;
  };
  bool gave_up = false;   // after one time-out this lane no longer waits: the launch ends quickly and the host sees *err
  auto finish = [&](int kind, const unsigned char *slot, LLWord w, unsigned tag) GIRIH_LAMBDA_INLINE -> R {
    if (kind == 1 && (w.t0 != tag || w.t1 != tag)) {   // rare: the neighbour is behind
      unsigned spins = 0;
#pragma unroll 1
      while ((w.t0 != tag || w.t1 != tag) && !gave_up) {
        if (++spins > R1X_SPIN_LIMIT) { gave_up = true; *a.err = 1; break; }
        spin_pause();   // a short sleep: a warp that re-polls at once floods the slot's L2 line and delays the store it waits for
        w = ll_load(slot);
      }
    }
    return ll_value(w);
  };

  auto body = [&](auto phase_tag, auto frame_tag, const int it) GIRIH_LAMBDA_INLINE {
    constexpr int PH = decltype(phase_tag)::value;
    constexpr bool FRAME = decltype(frame_tag)::value;
    constexpr int iB = PH, iC = (PH + 1) % 3, iF = (PH + 2) % 3;
    constexpr int RING_CUR = PH, RING_PREV = (PH + 2) % 3;      // ring slot of iteration it / it - 1 (it % 3 == PH)
    const int zin = zb - T + it;
    const int cur = it & 1;
    const unsigned tag = a.seq0 + (unsigned)it + 1u;           // this iteration's tag; the previous one is tag - 1

    auto publish = [&](auto level_tag, const R (&P)[PY][VX]) {
      // rim of the newest plane of level l: to the other warps through shared memory, to the other tiles through L2
      constexpr int l = decltype(level_tag)::value;
      st128<R>(edge_ptr(l, cur, warp, 0) + lane * VX, P[0]);
      st128<R>(edge_ptr(l, cur, warp, 1) + lane * VX, P[PY - 1]);
      if constexpr (((XDBG & 1) != 0 && (XDBG & 8) == 0) || l == 0) return;   // level 0 is read from the input array by everybody
#pragma unroll
      for (int j = 0; j < PY; ++j) {
        sts_if(lane_first, xo + j, P[j][0]);
        sts_if(lane_last, xo + PY + j, P[j][VX - 1]);
      }
    };
    // start of stage l >= 1: the rim of level l (parked in xo by the previous stage, __syncwarp in between) leaves
    auto send_x = [&](auto level_tag) GIRIH_LAMBDA_INLINE {
      constexpr int l = decltype(level_tag)::value;
      constexpr int off = (RING_CUR * RING_SLOTS + l * LEVEL_SLOTS) * 16;
      if constexpr ((XDBG & 1) != 0 || l == 0) return;
      R v = (R)0;
      lds_if(xpub8 != nullptr, xo + lane, v);
      ll_store_if<off>(xpub8, v, tag);
    };
    // the rows above / below the tile leave with the first / last warp: a warp-uniform branch, kept out of the hot block
    auto publish_y = [&](auto level_tag, const R (&P)[PY][VX]) GIRIH_LAMBDA_INLINE {
      constexpr int l = decltype(level_tag)::value;
      constexpr int off = (RING_CUR * RING_SLOTS + l * LEVEL_SLOTS) * 16;
      if constexpr ((XDBG & 1) != 0 || l == 0) return;
      if (ypub != nullptr) {
        if (warp == 0) {
          ll_store_if<off>(ypub, P[0][0], tag);
          ll_store_if<off + 16>(ypub, P[0][VX - 1], tag);
        } else {
          ll_store_if<off>(ypub, P[PY - 1][0], tag);
          ll_store_if<off + 16>(ypub, P[PY - 1][VX - 1], tag);
        }
      }
    };

    publish(Level<0>{}, S[0][iF]);
    publish_y(Level<0>{}, S[0][iF]);

    R Ofin[PY][VX];
    auto level = [&](auto level_tag) GIRIH_LAMBDA_INLINE {
      constexpr int l = decltype(level_tag)::value;
      const int zc = zin - l - 1;
      R (&Bp)[PY][VX] = S[l][iB];
      R (&Cp)[PY][VX] = S[l][iC];
      R (&Fp)[PY][VX] = S[l][iF];

      // ---- this stage's poll round: issue now, finish after the arithmetic (see "polling" above) ---------------
      constexpr int lv = (l + 1) % T;                       // level collected in this stage
      constexpr bool PREV = (l + 1 < T);                    // ... of the previous iteration (else: level 0 of this one)
      constexpr int RING_P = PREV ? RING_PREV : RING_CUR;
      const unsigned ptag = PREV ? tag - 1u : tag;
      const int zp = PREV ? zin - 1 - lv : zin;             // plane of that level's newest values
      const bool pactive = !PREV || it > 0;
      const unsigned char *const px0 = inb + (size_t)(RING_P * RING_SLOTS + lv * LEVEL_SLOTS) * 16 + sX;
      const unsigned char *const py0 = inb + (size_t)(RING_P * RING_SLOTS + lv * LEVEL_SLOTS) * 16 + sY;
      // level 0 (PREV == false): the cell of the input array, whatever the tile's position
      const int kx = (XDBG & 2) ? 0 : (PREV ? (pactive ? kX : 0) : (f0X ? 2 : 0));
      const int ky = ((XDBG & 2) || !edge_warp) ? 0 : (PREV ? (pactive ? kY : 0) : (f0Y ? 2 : 0));
      const unsigned ox = PREV ? sX : fX, oy = PREV ? sY : fY;
      send_x(Level<l>{});
      LLWord wx = issue(kx, px0, ox, zp, ptag);
      LLWord wy0 = issue(ky, py0, oy, zp, ptag);            // predicated off everywhere but in the first / last warp
      LLWord wy1 = issue(ky, py0 + 16, oy + 1u, zp, ptag);

      // rows just outside my strip: neighbouring warps, or (first / last warp) the neighbouring tiles
      R up[VX], dn[VX];
      {
        const int wu = (warp > 0) ? warp - 1 : NW, wd = (warp < NW - 1) ? warp + 1 : NW + 1;
        ld128s<R>(edge_ptr(l, cur ^ 1, wu, 1) + lane * VX, up);
        ld128s<R>(edge_ptr(l, cur ^ 1, wd, 0) + lane * VX, dn);
      }
      // columns just outside my span: lane 0 takes the left one, lane 31 the right one
      const R *const xq = xs + (l * 2 + (lane == 31 ? 1 : 0)) * PY;

      auto stage = [&](R (&O)[PY][VX]) {
#pragma unroll
        for (int j = 0; j < PY; ++j) {
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
          R left = __shfl_up_sync(0xffffffffu, Cp[j][VX - 1], 1);
          R right = __shfl_down_sync(0xffffffffu, Cp[j][0], 1);
          if constexpr ((XDBG & 4) == 0) {   // lane 0 / 31: the neighbour tile's column instead of the shuffle's own value
            lds_if(lane_first, xq + j, left);
            lds_if(lane_last, xq + j, right);
          }
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
        if constexpr (FRAME) {
          const unsigned upd = ((zc >= g.zlo) && (zc < g.zhi)) ? interior_xy : 0u;
#pragma unroll
          for (int j = 0; j < PY; ++j)
#pragma unroll
            for (int e = 0; e < VX; ++e) O[j][e] = ((upd >> (j * VX + e)) & 1u) ? O[j][e] : Cp[j][e];
        }
        if constexpr (l + 1 < T) publish(Level<l + 1>{}, O);
      };
      if constexpr (l + 1 < T) stage(S[l + 1][iF]);
      else stage(Ofin);
      if constexpr (l == 0) load_plane(zin + 1, S[0][iB], it + 1 < nit);
      if constexpr (l + 1 < T) publish_y(Level<l + 1>{}, S[l + 1][iF]);

      // ---- finish this stage's poll round: payloads into shared memory, for the next stage of this warp --------
#ifdef GIRIH_R1X_TRACE
      const bool tr = a.trace != nullptr && lane == 0 && (warp == 0 || warp == 3) && it >= 96 && it < 112;
      long long *const trp = a.trace + ((((cta * 2 + (warp == 0 ? 0 : 1)) * 16 + (it - 96)) * T + l) * 2);
      if (tr) trp[0] = clock64();
#endif
      if (kx != 0) xs[lv * 2 * PY + dX] = finish(kx, px0, wx, ptag);
      if (edge_warp) {
        if (ky != 0) {
          R t[VX];
          t[0] = finish(ky, py0, wy0, ptag);
          t[1] = finish(ky, py0 + 16, wy1, ptag);
          // consumed through parity cur ^ 1 of the iteration that reads it: this one (PREV) or the next one
          st128<R>(edge + dY + lv * EDGE_LVL + (PREV ? (cur ^ 1) : cur) * EDGE_PAR, t);
        }
      }
      __syncwarp();
#ifdef GIRIH_R1X_TRACE
      if (tr) trp[1] = clock64();
#endif
    };
    static_for<0, T>(level);

    // Ofin is level T at plane zin - T: every interior point of the tile is stored
    {
      const int zo = zin - T;
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
            const unsigned m = (interior_xy >> (j * VX)) & ((1u << VX) - 1u);
#pragma unroll
            for (int e = 0; e < VX; ++e)
              if ((m >> e) & 1u) q[(long long)j * g.px + e] = Ofin[j][e];
          }
        }
      }
    }
    __syncwarp();      // lanes that had to wait for a slot rejoin here: the CTA barrier is reached by whole warps
    __syncthreads();
  };

  auto step = [&](auto phase_tag, const int it) GIRIH_LAMBDA_INLINE {
    const int zin = zb - T + it;
    if (warp_masked || (zin - T < g.zlo) || (zin > g.zhi)) body(phase_tag, FrameTag<true>{}, it);
    else body(phase_tag, FrameTag<false>{}, it);
  };
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

}  // namespace girih
