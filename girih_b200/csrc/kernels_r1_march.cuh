// kernels_r1_march.cuh -- single-step (T = 1) sweep of the radius-1 star operators (slots 1, 2, 3, 5):
// the kernel behind ts 0 "Spatial Blocking" / ts 1 "Halo-first" and behind every unfused pass.
//
// One HBM pass per time step is pure bandwidth work (2 .. 9 words per lattice update), so this
// kernel is built for bytes in flight and few instructions per byte, not for arithmetic:
//   * no shared memory and no block-level barrier: every warp is independent
//   * a lane owns VX = 16 B / sizeof(Real) consecutive x points of PY consecutive rows: all global
//     accesses are coalesced 128-bit loads/stores of whole 128-byte lines (rows are 128-byte aligned)
//   * the thread marches along z.  A ring of RB = 4 register planes holds z-1, z, z+1 and the plane
//     z+2 that is still in flight; the ring rotates by unrolling the z loop RB times, so the z
//     column costs no register moves and every point of v is fetched from HBM exactly once
//   * x neighbours come from warp shuffles; lanes 0 / 31 fetch the one element beyond the warp's
//     32*VX span (a cache hit: the neighbouring warp streams that line)
//   * y neighbours inside the PY-row strip are registers; the two rows bordering the strip are read
//     as 128-bit loads that hit L1/L2 (another warp streams them), one iteration ahead
//   * tiles do not overlap: nothing is computed twice; warps that lie completely inside the domain
//     run a version without any per-lane predicate
#pragma once
#include "common.cuh"
#include "stencil_expr.cuh"

namespace girih {

template <typename R> struct MarchArgs {
  DevGrid g;
  const R *__restrict__ in;
  R *__restrict__ out;
  const R *__restrict__ coef;
  long long coef_stride;
  ConstCoef<R> cc;
  int zb0, ze0, zchunk;
};

template <typename R> struct RegNbM {
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

template <int I> struct MPhase { static constexpr int value = I; };
template <bool B> struct MFull { static constexpr bool value = B; };

// NWY warps per CTA, PY rows per thread: CTA tile = (32*VX) x (NWY*PY)
template <int K, typename R, int PY, int NWY, bool FM = false>
__global__ void __launch_bounds__(32 * NWY)
k_r1_march(const MarchArgs<R> a) {
  constexpr int VX = Vec<R>::N, WX = 32 * VX, RB = 4;
  constexpr int NCA = KTraits<K>::NCA;
  static_assert(KTraits<K>::R == 1 && KTraits<K>::TO == 1, "radius-1, first-order-in-time only");
  const DevGrid &g = a.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = g.X0 + (int)blockIdx.x * WX + lane * VX;
  const int y0 = g.Y0 + ((int)blockIdx.y * NWY + warp) * PY;
  const int zb = a.zb0 + (int)blockIdx.z * a.zchunk;
  const int ze = min(zb + a.zchunk, a.ze0);
  if (y0 >= g.Y0 + g.ny) return;                      // whole warp outside (no barriers in this kernel)
  // lanes that hold an interior point or the frame column right of it must load (their neighbours
  // read them through shuffles); lanes further out stay idle.  Rows past the frame are not touched.
  const bool act = x < g.X0 + g.nx + g.r;
  const int rows = min(PY, g.Y0 + g.ny - y0);         // interior rows of my strip (warp-uniform)
  unsigned inter = 0;                                 // bit e: point e of a row is interior
#pragma unroll
  for (int e = 0; e < VX; ++e)
    if (x + e < g.X0 + g.nx) inter |= 1u << e;
  // the whole warp is interior and no lane needs a predicate
  const bool full = (rows == PY) && __all_sync(0xffffffffu, inter == (1u << VX) - 1u);
  const long long off = (long long)y0 * g.px + x;
  const bool edge_lane = (lane == 0) || (lane == 31 && x + VX < g.px);
  const int edge_off = (lane == 0) ? -1 : VX;        // element beyond the warp's span, per row

  R ring[RB][PY][VX];   // ring[(ph + i) % RB] = plane z-1+i, i = 0..3, in phase ph = (z - zb) % RB
  R near[2][2][VX];     // [parity][0 = row y0-1, 1 = row y0+rows][VX] of the centre plane
  R edge[2][PY];        // [parity][row]: element left of lane 0 / right of lane 31
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < PY; ++j)
#pragma unroll
      for (int e = 0; e < VX; ++e) ring[i][j][e] = (R)0;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
#pragma unroll
    for (int e = 0; e < VX; ++e) { near[b][0][e] = (R)0; near[b][1][e] = (R)0; }
#pragma unroll
    for (int j = 0; j < PY; ++j) edge[b][j] = (R)0;
  }

  auto sweep = [&](auto full_tag) {
    constexpr bool FULL = decltype(full_tag)::value;
    auto fetch_plane = [&](int z, R (&dst)[PY][VX]) {     // my strip of plane z (the HBM stream)
      const R *pz = a.in + off + (long long)z * g.pxy;
#pragma unroll
      for (int j = 0; j < PY; ++j)
        if (FULL || (act && j < rows)) ld128<R>(pz + (long long)j * g.px, dst[j]);
    };
    auto fetch_near = [&](int z, R (&nr)[2][VX], R (&ed)[PY]) {   // cache hits around plane z
      const R *pz = a.in + off + (long long)z * g.pxy;
      if (FULL || act) {
        ld128<R>(pz - g.px, nr[0]);
        ld128<R>(pz + (long long)rows * g.px, nr[1]);
      }
      if (edge_lane) {
#pragma unroll
        for (int j = 0; j < PY; ++j)
          if (FULL || j < rows) ed[j] = __ldg(pz + (long long)j * g.px + edge_off);
      }
    };

    // prologue: planes zb-1, zb, zb+1 and the near data of plane zb
    fetch_plane(zb - 1, ring[0]);
    fetch_plane(zb, ring[1]);
    fetch_plane(zb + 1, ring[2]);
    fetch_near(zb, near[0], edge[0]);

    auto body = [&](auto phase_tag, const int z) {
      constexpr int PH = decltype(phase_tag)::value;
      R (&zm)[PY][VX] = ring[PH % RB];
      R (&zc)[PY][VX] = ring[(PH + 1) % RB];
      R (&zp)[PY][VX] = ring[(PH + 2) % RB];
      R (&nr)[2][VX] = near[PH & 1];
      R (&ed)[PY] = edge[PH & 1];
      // keep the stream two planes ahead, the cache-hit data one plane ahead
      if (z + 2 <= ze) fetch_plane(z + 2, ring[(PH + 3) % RB]);
      if (z + 1 < ze) fetch_near(z + 1, near[(PH + 1) & 1], edge[(PH + 1) & 1]);

      R *qz = a.out + off + (long long)z * g.pxy;
#pragma unroll
      for (int j = 0; j < PY; ++j) {
        R cf[NCA > 0 ? NCA : 1][VX];
        if constexpr (NCA > 0) {
          const R *cp = a.coef + off + (long long)z * g.pxy + (long long)j * g.px;
#pragma unroll
          for (int m = 0; m < NCA; ++m) {
#pragma unroll
            for (int e = 0; e < VX; ++e) cf[m][e] = (R)0;
            if (FULL || (act && j < rows)) ld128<R>(cp + (long long)m * a.coef_stride, cf[m]);
          }
        }
        R left = __shfl_up_sync(0xffffffffu, zc[j][VX - 1], 1);
        R right = __shfl_down_sync(0xffffffffu, zc[j][0], 1);
        if (lane == 0) left = ed[j];
        if (lane == 31) right = ed[j];
        R o[VX];
#pragma unroll
        for (int e = 0; e < VX; ++e) {
          RegNbM<R> n;
          n.c = zc[j][e];
          n.xm = (e > 0) ? zc[j][e > 0 ? e - 1 : 0] : left;
          n.xp = (e < VX - 1) ? zc[j][e < VX - 1 ? e + 1 : 0] : right;
          n.ym = (j > 0) ? zc[j > 0 ? j - 1 : 0][e] : nr[0][e];
          if constexpr (FULL) n.yp = (j < PY - 1) ? zc[j < PY - 1 ? j + 1 : 0][e] : nr[1][e];
          else n.yp = (j + 1 < rows) ? zc[j < PY - 1 ? j + 1 : 0][e] : nr[1][e];
          n.zm = zm[j][e];
          n.zp = zp[j][e];
          if constexpr (NCA > 0) {
            RegCoef<R, KTraits<K>::NCA> rc;   // (spelled out: g++ 13 rejects the local constexpr here)
#pragma unroll
            for (int m = 0; m < NCA; ++m) rc.v[m] = cf[m][e];
            o[e] = StencilExpr<K>::template eval<R, FM>(n, rc, (R)0, (R)0);
          } else {
            o[e] = StencilExpr<K>::template eval<R, FM>(n, a.cc, (R)0, (R)0);
          }
        }
        if constexpr (FULL) {
          st128<R>(qz + (long long)j * g.px, o);
        } else if (j < rows) {
          if (inter == (1u << VX) - 1u) {
            st128<R>(qz + (long long)j * g.px, o);
          } else {
#pragma unroll
            for (int e = 0; e < VX; ++e)
              if ((inter >> e) & 1u) qz[(long long)j * g.px + e] = o[e];
          }
        }
      }
    };

    int z = zb;
    for (; z + RB <= ze; z += RB) {
      body(MPhase<0>{}, z);
      body(MPhase<1>{}, z + 1);
      body(MPhase<2>{}, z + 2);
      body(MPhase<3>{}, z + 3);
    }
    if (z < ze) { body(MPhase<0>{}, z); ++z; }
    if (z < ze) { body(MPhase<1>{}, z); ++z; }
    if (z < ze) { body(MPhase<2>{}, z); }
  };
  if (full) sweep(MFull<true>{});
  else sweep(MFull<false>{});
}

}  // namespace girih
