// kernels_box.cuh -- single-step sweep of the 27-point box operator (table slot 7,
// src/kernels/stencils.c:227-242): the z-marching counterpart of k_r1_march for a stencil that also
// reads the xy-, xz-, yz-diagonal and corner neighbours.
//
//   * one warp = one row of 32*VX points, a lane owns VX consecutive x points (128-bit accesses)
//   * the thread marches along z and keeps, for the planes z-1, z, z+1, its own VX points of the rows
//     y-1, y, y+1 in registers, together with the element left and right of them (from the
//     neighbouring lane by shuffle; lanes 0 / 31 load the one element beyond the warp's span)
//   * the rows y-1 and y+1 are loaded by this thread as well -- the warps above and below stream the
//     same lines, so two of the three 128-bit loads per plane are L1/L2 hits and HBM traffic stays at one
//     read and one write of the grid
//   * a ring of RB = 4 planes (one in flight) rotates by unrolling the z loop, no register moves
//   * no shared memory, no barrier; tiles do not overlap
// 14 products and 26 sums per update in the reference's order (40 FP64 instructions, 27 with the
// "contract" option): at 512^3 fp64 the FP64 pipe and HBM need about the same time.
#pragma once
#include "common.cuh"
#include "kernels_r4.cuh"   // cp_async16 / cp_async_commit / cp_async_wait
#include "stencil_expr.cuh"

namespace girih {

template <typename R> struct BoxArgs {
  DevGrid g;
  const R *__restrict__ in;
  R *__restrict__ out;
  ConstCoef<R> cc;
  int zb0, ze0, zchunk;
};

// the 27 neighbours of one point, gathered from the register planes (pure renaming after unrolling)
template <typename R> struct BoxNb {
  R v[3][3][3];   // [dz+1][dy+1][dx+1]
  template <int DX, int DY, int DZ> __device__ __forceinline__ R at() const { return v[DZ + 1][DY + 1][DX + 1]; }
};

template <int I> struct BPhase { static constexpr int value = I; };

template <typename R, int NW, bool FM = false>
__global__ void __launch_bounds__(32 * NW, 16 / NW)   // 16 warps per SM: at most 128 registers
k_box_march(const BoxArgs<R> a) {
  constexpr int VX = Vec<R>::N, WX = 32 * VX, RB = 4;
  const DevGrid &g = a.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = g.X0 + (int)blockIdx.x * WX + lane * VX;
  const int y = g.Y0 + (int)blockIdx.y * NW + warp;
  const int zb = a.zb0 + (int)blockIdx.z * a.zchunk;
  const int ze = min(zb + a.zchunk, a.ze0);
  if (y >= g.Y0 + g.ny) return;                      // whole warp outside (no barriers in this kernel)
  const bool act = x < g.X0 + g.nx + g.r;            // interior lanes and the frame column right of them
  unsigned inter = 0;
#pragma unroll
  for (int e = 0; e < VX; ++e)
    if (x + e < g.X0 + g.nx) inter |= 1u << e;
  const long long off = (long long)(y - 1) * g.px + x;   // my points in row y-1
  const bool edge_lane = (lane == 0) || (lane == 31 && x + VX < g.px);
  const int edge_off = (lane == 0) ? -1 : VX;

  R ctr[RB][3][VX];   // ctr[(ph + i) % RB][row] = my points of plane z-1+i, rows y-1, y, y+1
  R lr[RB][3][2];     // element left / right of them
  R ed[RB][3];        // lanes 0 / 31: the element beyond the warp's span (loaded with the plane)
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int e = 0; e < VX; ++e) ctr[i][r][e] = (R)0;
      lr[i][r][0] = lr[i][r][1] = (R)0;
      ed[i][r] = (R)0;
    }

  auto fetch = [&](int z, R (&c)[3][VX], R (&d)[3]) {
    const R *pz = a.in + off + (long long)z * g.pxy;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (act) ld128<R>(pz + (long long)r * g.px, c[r]);
      if (edge_lane) d[r] = __ldg(pz + (long long)r * g.px + edge_off);
    }
  };
  auto sides = [&](const R (&c)[3][VX], const R (&d)[3], R (&s)[3][2]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      R left = __shfl_up_sync(0xffffffffu, c[r][VX - 1], 1);
      R right = __shfl_down_sync(0xffffffffu, c[r][0], 1);
      if (lane == 0) left = d[r];
      if (lane == 31) right = d[r];
      s[r][0] = left;
      s[r][1] = right;
    }
  };

  // prologue: planes zb-1, zb complete with their sides, plane zb+1 in flight
  fetch(zb - 1, ctr[0], ed[0]);
  fetch(zb, ctr[1], ed[1]);
  fetch(zb + 1, ctr[2], ed[2]);
  sides(ctr[0], ed[0], lr[0]);
  sides(ctr[1], ed[1], lr[1]);

  auto body = [&](auto phase_tag, const int z) {
    constexpr int PH = decltype(phase_tag)::value;
    constexpr int S0 = PH % RB, S1 = (PH + 1) % RB, S2 = (PH + 2) % RB, S3 = (PH + 3) % RB;
    if (z + 2 <= ze) fetch(z + 2, ctr[S3], ed[S3]);   // keep one plane in flight
    sides(ctr[S2], ed[S2], lr[S2]);                    // plane z+1 has arrived
    R o[VX];
#pragma unroll
    for (int e = 0; e < VX; ++e) {
      BoxNb<R> n;
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) {
        const int s = (dz == 0) ? S0 : (dz == 1) ? S1 : S2;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          n.v[dz][r][0] = (e > 0) ? ctr[s][r][e > 0 ? e - 1 : 0] : lr[s][r][0];
          n.v[dz][r][1] = ctr[s][r][e];
          n.v[dz][r][2] = (e < VX - 1) ? ctr[s][r][e < VX - 1 ? e + 1 : 0] : lr[s][r][1];
        }
      }
      o[e] = StencilExpr<7>::template eval<R, FM>(n, a.cc, (R)0, (R)0);
    }
    R *q = a.out + off + (long long)z * g.pxy + g.px;   // row y
    if (inter == (1u << VX) - 1u) {
      st128<R>(q, o);
    } else {
#pragma unroll
      for (int e = 0; e < VX; ++e)
        if ((inter >> e) & 1u) q[e] = o[e];
    }
  };

  int z = zb;
  for (; z + RB <= ze; z += RB) {
    body(BPhase<0>{}, z);
    body(BPhase<1>{}, z + 1);
    body(BPhase<2>{}, z + 2);
    body(BPhase<3>{}, z + 3);
  }
  if (z < ze) { body(BPhase<0>{}, z); ++z; }
  if (z < ze) { body(BPhase<1>{}, z); ++z; }
  if (z < ze) { body(BPhase<2>{}, z); }
}

// ------------------------------------------------------------------------------------------------------
// k_box_async (round 2) -- the same operator with the plane stream prefetched by cp.async into a 4-stage
// shared-memory ring, the schedule of k_r4_async.  k_box_march keeps ONE plane in flight per thread and reads
// every row three times through L1: 8 KB of new bytes in flight per SM, spills of in-flight load destinations at
// the 128-register cap, 40-47% of the HBM roofline (ncu: long scoreboard 5 cycles per issue).  Here
//   * the CTA tile is WX x NW points (one row per warp); every plane of the tile, with its one-point rim in x
//     and y, is staged in shared memory: one 16-byte cp.async per thread for its own points, one more for a
//     rim item (rows -1 / NW as vectors, the rim columns as single elements), one commit group per plane,
//     THREE planes in flight at no register cost (24 KB per CTA in fp64)
//   * a plane is read from shared memory ONCE, when it becomes plane z+1 of the thread's column: its rows
//     y-1, y, y+1 with the element left and right of them (9 loads) go into the register ring that already
//     holds the planes z-1 and z
//   * one __syncthreads per plane; tiles do not overlap
// ------------------------------------------------------------------------------------------------------
#ifndef GIRIH_CUDA_EMU
template <int BYTES> __device__ __forceinline__ void cp_async_small(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(s), "l"(gmem), "n"(BYTES) : "memory");
}
#endif   // the test suite's CPU SIMT emulator supplies its own version

template <typename R, int NW> struct BoxACfg {
  static constexpr int VX = Vec<R>::N;
  static constexpr int WX = 32 * VX;
  static constexpr int NT = 32 * NW;
  static constexpr int NS = 4;                       // stages: the plane being read + 3 in flight
  static constexpr int PAD = VX;                     // the rim column left of the tile sits at PAD - 1 (rows stay 16-byte aligned)
  static constexpr int SP = WX + 2 * PAD;            // shared row pitch (elements)
  static constexpr int SROWS = NW + 2;
  static constexpr int PLANE = SROWS * SP;
  static constexpr int NRIM = 2 * 32 + 2 * SROWS;    // rim items of a plane: 2 rows of 32 vectors, 2 columns of SROWS elements
  static constexpr size_t SMEM = (size_t)NS * PLANE * sizeof(R);
  static_assert(NRIM <= NT, "one rim item per thread");
};

template <typename R, int NW, bool FM = false>
__global__ void __launch_bounds__(32 * NW)
k_box_async(const BoxArgs<R> a) {
  using Cfg = BoxACfg<R, NW>;
  constexpr int VX = Cfg::VX, WX = Cfg::WX, NS = Cfg::NS, PAD = Cfg::PAD, SP = Cfg::SP, PLANE = Cfg::PLANE;
  constexpr int SROWS = Cfg::SROWS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *planes = reinterpret_cast<R *>(smem_raw);       // [NS][SROWS][SP]

  const DevGrid &g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0t = g.X0 + (int)blockIdx.x * WX, y0t = g.Y0 + (int)blockIdx.y * NW;
  const int x = x0t + lane * VX, y = y0t + warp;
  const int zb = a.zb0 + (int)blockIdx.z * a.zchunk;
  const int ze = min(zb + a.zchunk, a.ze0);
  const bool ok = (x + VX <= g.px) && (y < g.ny_dev);        // my vector may be loaded
  unsigned inter = 0;
#pragma unroll
  for (int e = 0; e < VX; ++e)
    if ((x + e < g.X0 + g.nx) && (y < g.Y0 + g.ny)) inter |= 1u << e;
  const long long off = (long long)y * g.px + x;

  // my rim item of a plane: tid < 64: a vector of row -1 / NW; then the elements left / right of rows -1 .. NW
  int rim_s = 0, rim_kind = 0;       // 0 none, 1 vector, 2 single element
  long long rim_g = 0;
  if (tid < 64) {
    const int srow = (tid < 32) ? 0 : SROWS - 1, vv = tid & 31;
    const int gx = x0t + vv * VX, gy = y0t - 1 + srow;
    rim_s = srow * SP + PAD + vv * VX;
    rim_g = (long long)gy * g.px + gx;
    rim_kind = (gx + VX <= g.px && gy >= 0 && gy < g.ny_dev) ? 1 : 0;
  } else if (tid < Cfg::NRIM) {
    const int k = tid - 64, srow = k >> 1, side = k & 1;
    const int gx = side == 0 ? x0t - 1 : x0t + WX, gy = y0t - 1 + srow;
    rim_s = srow * SP + (side == 0 ? PAD - 1 : PAD + WX);
    rim_g = (long long)gy * g.px + gx;
    rim_kind = (gx >= 0 && gx < g.px && gy >= 0 && gy < g.ny_dev) ? 2 : 0;
  }
  // G(p): plane p of the tile and its rim, one commit group (possibly empty past the chunk)
  auto issue_group = [&](int p) {
    if (p <= ze && p >= 0 && p < g.nz_dev) {
      R *buf = planes + (size_t)((p - zb) & (NS - 1)) * PLANE;
      const R *pv = a.in + (long long)p * g.pxy;
      if (ok) cp_async16(buf + (1 + warp) * SP + PAD + lane * VX, pv + off);
      if (rim_kind == 1) cp_async16(buf + rim_s, pv + rim_g);
      else if (rim_kind == 2) cp_async_small<(int)sizeof(R)>(buf + rim_s, pv + rim_g);
    }
    cp_async_commit();
  };

  R ctr[3][3][VX];   // ctr[(ph + i) % 3][row] = my points of plane z-1+i, rows y-1, y, y+1
  R lr[3][3][2];     // element left / right of them
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int e = 0; e < VX; ++e) ctr[i][r][e] = (R)0;
      lr[i][r][0] = lr[i][r][1] = (R)0;
    }
  // a staged plane -> my three rows and their side elements
  auto take = [&](const R *s, R (&c)[3][VX], R (&d)[3][2]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const R *row = s + (warp + r) * SP + PAD + lane * VX;
      ld128s<R>(row, c[r]);
      d[r][0] = row[-1];
      d[r][1] = row[VX];
    }
  };
  static_assert(NS == 4, "stage index uses a mask");
  // prologue: planes zb-1 and zb go through the ring too (stages 3 and 0); then zb+1, zb+2, zb+3 are in flight
  issue_group(zb - 1);
  issue_group(zb);
  issue_group(zb + 1);
  issue_group(zb + 2);
  cp_async_wait<2>();
  __syncthreads();
  take(planes + (size_t)(NS - 1) * PLANE, ctr[0], lr[0]);
  take(planes + (size_t)0 * PLANE, ctr[1], lr[1]);
  __syncthreads();                                   // everyone has taken plane zb-1: its stage may be refilled
  issue_group(zb + 3);

  auto body = [&](auto phase_tag, const int z) {
    constexpr int PH = decltype(phase_tag)::value;
    constexpr int S0 = PH % 3, S1 = (PH + 1) % 3, S2 = (PH + 2) % 3;
    cp_async_wait<2>();                              // G(z+1) has landed (for this thread); z+2, z+3 may be in flight
    __syncthreads();                                 // ... for everyone; and everyone has left iteration z-1
    issue_group(z + 4);                              // refills the stage of plane z, taken in iteration z-1
    take(planes + (size_t)((z + 1 - zb) & (NS - 1)) * PLANE, ctr[S2], lr[S2]);
    R o[VX];
#pragma unroll
    for (int e = 0; e < VX; ++e) {
      BoxNb<R> n;
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) {
        const int s = (dz == 0) ? S0 : (dz == 1) ? S1 : S2;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          n.v[dz][r][0] = (e > 0) ? ctr[s][r][e > 0 ? e - 1 : 0] : lr[s][r][0];
          n.v[dz][r][1] = ctr[s][r][e];
          n.v[dz][r][2] = (e < VX - 1) ? ctr[s][r][e < VX - 1 ? e + 1 : 0] : lr[s][r][1];
        }
      }
      o[e] = StencilExpr<7>::template eval<R, FM>(n, a.cc, (R)0, (R)0);
    }
    R *q = a.out + off + (long long)z * g.pxy;
    if (inter == (1u << VX) - 1u) {
      st128<R>(q, o);
    } else if (inter != 0u) {
#pragma unroll
      for (int e = 0; e < VX; ++e)
        if ((inter >> e) & 1u) q[e] = o[e];
    }
  };

  int z = zb;
  for (; z + 3 <= ze; z += 3) {
    body(BPhase<0>{}, z);
    body(BPhase<1>{}, z + 1);
    body(BPhase<2>{}, z + 2);
  }
  if (z < ze) { body(BPhase<0>{}, z); ++z; }
  if (z < ze) { body(BPhase<1>{}, z); }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------------
// k_box_lean: k_box_async with the per-plane overhead taken out.  In k_box_async 95 of the 175 instructions of a
// plane iteration (fp64, two updates per thread) are not arithmetic: 64-bit multiplies for every plane address,
// three-way branches around the copies (reconvergence barriers included), three range compares per copy group.
// Here every address is a running pointer (one 64-bit add per plane), the stage index is a running counter, and the
// copies are PREDICATED instructions (@p cp.async) instead of branches, so a plane iteration is one basic block.
// Same staging, same arithmetic, same results.
// ------------------------------------------------------------------------------------------------------
#ifndef GIRIH_CUDA_EMU
__device__ __forceinline__ void cp_async16_if(void *smem, const void *gmem, bool p) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q cp.async.cg.shared.global [%0], [%1], 16;\n}\n" ::"r"(s), "l"(gmem),
               "r"((int)p)
               : "memory");
}
template <int BYTES> __device__ __forceinline__ void cp_async_small_if(void *smem, const void *gmem, bool p) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q cp.async.ca.shared.global [%0], [%1], %3;\n}\n" ::"r"(s), "l"(gmem),
               "r"((int)p), "n"(BYTES)
               : "memory");
}
#else
static inline void cp_async16_if(void *smem, const void *gmem, bool p) { if (p) cuda_emu::cp_async_issue(smem, gmem); }
template <int BYTES> static inline void cp_async_small_if(void *smem, const void *gmem, bool p) {
  if (p) cuda_emu::cp_async_issue(smem, gmem, BYTES);
}
#endif

template <typename R, int NW, bool FM = false>
__global__ void __launch_bounds__(32 * NW)
k_box_lean(const BoxArgs<R> a) {
  using Cfg = BoxACfg<R, NW>;
  constexpr int VX = Cfg::VX, WX = Cfg::WX, NS = Cfg::NS, PAD = Cfg::PAD, SP = Cfg::SP, PLANE = Cfg::PLANE;
  constexpr int SROWS = Cfg::SROWS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *planes = reinterpret_cast<R *>(smem_raw);       // [NS][SROWS][SP]

  const DevGrid &g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0t = g.X0 + (int)blockIdx.x * WX, y0t = g.Y0 + (int)blockIdx.y * NW;
  const int x = x0t + lane * VX, y = y0t + warp;
  const int zb = a.zb0 + (int)blockIdx.z * a.zchunk;
  const int ze = min(zb + a.zchunk, a.ze0);
  const bool ok = (x + VX <= g.px) && (y < g.ny_dev);        // my vector may be loaded
  unsigned inter = 0;
#pragma unroll
  for (int e = 0; e < VX; ++e)
    if ((x + e < g.X0 + g.nx) && (y < g.Y0 + g.ny)) inter |= 1u << e;
  const long long off = (long long)y * g.px + x;

  // my rim item of a plane: tid < 64: a vector of row -1 / NW; then the elements left / right of rows -1 .. NW
  int rim_s = 0;
  bool rim_vec = false, rim_one = false;
  long long rim_g = 0;
  if (tid < 64) {
    const int srow = (tid < 32) ? 0 : SROWS - 1, vv = tid & 31;
    const int gx = x0t + vv * VX, gy = y0t - 1 + srow;
    rim_s = srow * SP + PAD + vv * VX;
    rim_g = (long long)gy * g.px + gx;
    rim_vec = (gx + VX <= g.px && gy >= 0 && gy < g.ny_dev);
  } else if (tid < Cfg::NRIM) {
    const int k = tid - 64, srow = k >> 1, side = k & 1;
    const int gx = side == 0 ? x0t - 1 : x0t + WX, gy = y0t - 1 + srow;
    rim_s = srow * SP + (side == 0 ? PAD - 1 : PAD + WX);
    rim_g = (long long)gy * g.px + gx;
    rim_one = (gx >= 0 && gx < g.px && gy >= 0 && gy < g.ny_dev);
  }
  if (!rim_vec && !rim_one) rim_g = off;             // never dereferenced, but keep the running pointer inside the array
  // the copy group of the next plane: running pointers, running stage
  const int p_lim = min(ze + 1, g.nz_dev);           // planes zb-1 .. p_lim-1 are staged, later groups are empty
  int p_next = zb - 1, st_next = NS - 1;
  const R *gp_my = a.in + off + (long long)(zb - 1) * g.pxy;
  const R *gp_rim = a.in + rim_g + (long long)(zb - 1) * g.pxy;
  const int my_s = (1 + warp) * SP + PAD + lane * VX;
  auto issue_next = [&]() {
    const bool live = (p_next < p_lim) && (p_next >= 0);
    R *buf = planes + st_next * PLANE;
    cp_async16_if(buf + my_s, gp_my, live && ok);
    cp_async16_if(buf + rim_s, gp_rim, live && rim_vec);
    cp_async_small_if<(int)sizeof(R)>(buf + rim_s, gp_rim, live && rim_one);
    cp_async_commit();
    gp_my += g.pxy;
    gp_rim += g.pxy;
    ++p_next;
    st_next = (st_next + 1) & (NS - 1);
  };

  R ctr[3][3][VX];   // ctr[(ph + i) % 3][row] = my points of plane z-1+i, rows y-1, y, y+1
  R lr[3][3][2];     // element left / right of them
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int e = 0; e < VX; ++e) ctr[i][r][e] = (R)0;
      lr[i][r][0] = lr[i][r][1] = (R)0;
    }
  const int tk = warp * SP + PAD + lane * VX;        // my rows y-1, y, y+1 start here, SP apart
  int st_take = NS - 1;                              // stage of the next plane to take (plane zb-1 first)
  auto take_next = [&](R (&c)[3][VX], R (&d)[3][2]) {
    const R *s = planes + st_take * PLANE + tk;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const R *row = s + r * SP;
      ld128s<R>(row, c[r]);
      d[r][0] = row[-1];
      d[r][1] = row[VX];
    }
    st_take = (st_take + 1) & (NS - 1);
  };
  static_assert(NS == 4, "stage index uses a mask");
  // prologue: planes zb-1 and zb go through the ring too (stages 3 and 0); then zb+1, zb+2, zb+3 are in flight
  issue_next();
  issue_next();
  issue_next();
  issue_next();
  cp_async_wait<2>();
  __syncthreads();
  take_next(ctr[0], lr[0]);
  take_next(ctr[1], lr[1]);
  __syncthreads();                                   // everyone has taken plane zb-1: its stage may be refilled
  issue_next();

  R *q = a.out + off + (long long)zb * g.pxy;
  const bool full = inter == (1u << VX) - 1u;
  auto body = [&](auto phase_tag) {
    constexpr int PH = decltype(phase_tag)::value;
    constexpr int S0 = PH % 3, S1 = (PH + 1) % 3, S2 = (PH + 2) % 3;
    cp_async_wait<2>();                              // G(z+1) has landed (for this thread); z+2, z+3 may be in flight
    __syncthreads();                                 // ... for everyone; and everyone has left iteration z-1
    issue_next();                                    // G(z+4): refills the stage of plane z, taken in iteration z-1
    take_next(ctr[S2], lr[S2]);
    R o[VX];
#pragma unroll
    for (int e = 0; e < VX; ++e) {
      BoxNb<R> n;
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) {
        const int s = (dz == 0) ? S0 : (dz == 1) ? S1 : S2;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          n.v[dz][r][0] = (e > 0) ? ctr[s][r][e > 0 ? e - 1 : 0] : lr[s][r][0];
          n.v[dz][r][1] = ctr[s][r][e];
          n.v[dz][r][2] = (e < VX - 1) ? ctr[s][r][e < VX - 1 ? e + 1 : 0] : lr[s][r][1];
        }
      }
      o[e] = StencilExpr<7>::template eval<R, FM>(n, a.cc, (R)0, (R)0);
    }
    if (full) {
      st128<R>(q, o);
    } else if (inter != 0u) {
#pragma unroll
      for (int e = 0; e < VX; ++e)
        if ((inter >> e) & 1u) q[e] = o[e];
    }
    q += g.pxy;
  };

  int z = zb;
  for (; z + 3 <= ze; z += 3) {
    body(BPhase<0>{});
    body(BPhase<1>{});
    body(BPhase<2>{});
  }
  if (z < ze) { body(BPhase<0>{}); ++z; }
  if (z < ze) { body(BPhase<1>{}); }
  cp_async_wait<0>();
}

}  // namespace girih
