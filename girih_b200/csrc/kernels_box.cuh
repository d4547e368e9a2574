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

}  // namespace girih
