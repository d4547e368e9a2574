// kernels_r4.cuh -- z-streamed single-step sweep for the radius-4 (25-point) star operators:
// slot 0 (constant coefficients, 2nd order in time: reads v, u, roc2, writes u) and slot 4
// (13 per-point axis-symmetric coefficient arrays, 1st order in time).
//
// Schedule (one CTA = NW warps, one row per warp: tile WX x NW, marching along z):
//   * a lane owns VX = 16 B / sizeof(Real) consecutive x points: 128-bit coalesced global access
//   * the z column lives in a ring of RB = 10 register vectors (planes z-4 .. z+4 plus the plane
//     z+5 that is still in flight).  The z loop is unrolled RB times so the ring rotates by renaming,
//     without register moves; each plane of v is read from HBM exactly once per tile
//   * the centre plane z is staged in shared memory together with its 4-wide x/y halo strips
//     (double buffered, ONE __syncthreads per plane); x and y neighbours are read back as 128-bit
//     row/column windows
//   * halo strips, u(old) and roc2 of the NEXT plane are fetched one iteration ahead into registers
//   * per-point coefficients (slot 4) are touched at the thread's own points only and stream
//     straight from HBM into registers
// No temporal fusion here: with r = 4 the overlapped tile of a fused sweep wastes more than half
// of an SM-sized tile (halo 2*T*4 per axis) and slot 0 would need two extra arrays because its
// update reads the level it overwrites.  DESIGN.md, "r = 4 operators".
#pragma once
#include "common.cuh"
#include "stencil_expr.cuh"

namespace girih {

template <typename R> struct R4Args {
  DevGrid g;
  const R *__restrict__ v;      // newest level (read with halos)
  R *__restrict__ u;            // slot 0: level before v on input, new level on output; slot 4: output
  const R *__restrict__ roc2;   // slot 0
  const R *__restrict__ coef;   // slot 4
  long long coef_stride;
  ConstCoef<R> cc;              // slot 0
  int zb0, ze0, zchunk;
};

template <typename R, int NW> struct R4Cfg {
  static constexpr int RAD = 4;
  static constexpr int VX = Vec<R>::N;
  static constexpr int WX = 32 * VX;
  static constexpr int H = NW;
  static constexpr int NT = 32 * NW;
  static constexpr int RB = 10;
  static constexpr int SP = WX + 2 * RAD;        // shared row pitch (elements), 16-byte multiple
  static constexpr int SROWS = H + 2 * RAD;
  static constexpr int HXV = RAD / VX;           // halo vectors per row side
  static constexpr int NHV = 2 * RAD * 32 + H * 2 * HXV;   // halo vectors per plane
  static constexpr int HPT = (NHV + NT - 1) / NT;           // halo vectors per thread
  static constexpr size_t SMEM = (size_t)2 * SROWS * SP * sizeof(R);
  static_assert(RAD % VX == 0, "halo must be whole vectors");
};

// neighbour accessor: z from the register ring, x from the row window, y from the column window
template <typename R, int PH> struct RegNb4 {
  static constexpr int VX = Vec<R>::N;
  const R (*ring)[VX];       // ring[(PH + i) % 10] = plane z-4+i
  const R *xr;               // x-4 .. x+VX+3
  const R (*yc)[VX];         // rows y-4 .. y+4 (index 4 = my row)
  int e;
  template <int DX, int DY, int DZ> __device__ __forceinline__ R at() const {
    if constexpr (DZ != 0) return ring[(PH + 4 + DZ) % 10][e];
    else if constexpr (DX != 0) return xr[4 + e + DX];
    else if constexpr (DY != 0) return yc[4 + DY][e];
    else return ring[(PH + 4) % 10][e];
  }
};

template <int I> struct R4Phase { static constexpr int value = I; };

template <int K, typename R, int NW, bool FM = false>
__global__ void __launch_bounds__(32 * NW)
k_r4(const R4Args<R> a) {
  using Cfg = R4Cfg<R, NW>;
  constexpr int RAD = Cfg::RAD, VX = Cfg::VX, WX = Cfg::WX, H = Cfg::H, NT = Cfg::NT, RB = Cfg::RB;
  constexpr int SP = Cfg::SP, SROWS = Cfg::SROWS, HXV = Cfg::HXV, NHV = Cfg::NHV, HPT = Cfg::HPT;
  constexpr int NCA = KTraits<K>::NCA;
  constexpr bool TO2 = KTraits<K>::TO == 2;
  static_assert(KTraits<K>::R == 4, "radius-4 operators only");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *sm = reinterpret_cast<R *>(smem_raw);   // [2][SROWS][SP]

  const DevGrid &g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0t = g.X0 + (int)blockIdx.x * WX, y0t = g.Y0 + (int)blockIdx.y * H;
  const int x = x0t + lane * VX, y = y0t + warp;
  const int zb = a.zb0 + (int)blockIdx.z * a.zchunk;
  const int ze = min(zb + a.zchunk, a.ze0);

  const bool ok = (x + VX <= g.px) && (y < g.ny_dev);     // my vector may be loaded
  unsigned inter = 0;                                      // bit e: point is interior in x and y
#pragma unroll
  for (int e = 0; e < VX; ++e)
    if ((x + e < g.X0 + g.nx) && (y < g.Y0 + g.ny)) inter |= 1u << e;
  const long long off = (long long)y * g.px + x;

  // my share of the halo strips of a plane: top/bottom 4 rows over the tile width, left/right 4
  // columns over the tile height (a star stencil never reads the corners)
  int hs_off[HPT];
  long long hg_off[HPT];
  bool h_ok[HPT];
#pragma unroll
  for (int h = 0; h < HPT; ++h) {
    const int item = tid + h * NT;
    int srow, scol;
    if (item < 2 * RAD * 32) {
      const int rr = item >> 5, vv = item & 31;
      srow = (rr < RAD) ? rr : H + rr;
      scol = RAD + vv * VX;
    } else {
      const int it2 = item - 2 * RAD * 32;
      const int row = it2 / (2 * HXV), k = it2 % (2 * HXV);
      const int side = k / HXV, hv = k % HXV;
      srow = RAD + row;
      scol = (side == 0) ? hv * VX : RAD + WX + hv * VX;
    }
    const int gx = x0t - RAD + scol, gy = y0t - RAD + srow;
    hs_off[h] = srow * SP + scol;
    hg_off[h] = (long long)gy * g.px + gx;
    h_ok[h] = (item < NHV) && (gx >= 0) && (gx + VX <= g.px) && (gy >= 0) && (gy < g.ny_dev);
  }

  R ring[RB][VX];           // ring[(ph + i) % RB] = plane z-4+i (i = 0..8), (ph + 9) % RB in flight
  R hal[2][HPT][VX];        // [parity] halo vectors of the plane staged in that parity
  R uo[2][VX], rc[2][VX];   // [parity] u(old) and roc2 of that plane (slot 0)
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int e = 0; e < VX; ++e) ring[i][e] = (R)0;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
#pragma unroll
    for (int h = 0; h < HPT; ++h)
#pragma unroll
      for (int e = 0; e < VX; ++e) hal[b][h][e] = (R)0;
#pragma unroll
    for (int e = 0; e < VX; ++e) { uo[b][e] = (R)0; rc[b][e] = (R)0; }
  }

  auto fetch_v = [&](int z, R (&dst)[VX]) {
    if (ok && z >= 0 && z < g.nz_dev) ld128<R>(a.v + off + (long long)z * g.pxy, dst);
  };
  auto fetch_side = [&](int z, R (&hd)[HPT][VX], R (&ud)[VX], R (&rd)[VX]) {
    if (z >= 0 && z < g.nz_dev) {
      const R *p = a.v + (long long)z * g.pxy;
#pragma unroll
      for (int h = 0; h < HPT; ++h)
        if (h_ok[h]) ld128<R>(p + hg_off[h], hd[h]);
      if constexpr (TO2) {
        if (ok) {
          ld128g<R>(a.u + off + (long long)z * g.pxy, ud);       // same kernel overwrites u: coherent load
          ld128<R>(a.roc2 + off + (long long)z * g.pxy, rd);
        }
      }
    }
  };

  // prologue: planes zb-4 .. zb+4 of my column, side data of plane zb
#pragma unroll
  for (int i = 0; i < 9; ++i) fetch_v(zb - 4 + i, ring[i]);
  fetch_side(zb, hal[0], uo[0], rc[0]);

  auto body = [&](auto phase_tag, const int z) {
    constexpr int PH = decltype(phase_tag)::value;
    constexpr int PAR = PH & 1;
    R *s = sm + (size_t)PAR * SROWS * SP;
    // keep the column one plane ahead of what the next iteration needs, side data one plane ahead
    if (z + 1 < ze) {
      fetch_v(z + 5, ring[(PH + 9) % RB]);
      fetch_side(z + 1, hal[PAR ^ 1], uo[PAR ^ 1], rc[PAR ^ 1]);
    }
    // stage plane z: my own points from the register ring, my share of the halo strips
    st128<R>(s + (RAD + warp) * SP + RAD + lane * VX, ring[(PH + 4) % RB]);
#pragma unroll
    for (int h = 0; h < HPT; ++h)
      if (tid + h * NT < NHV) st128<R>(s + hs_off[h], hal[PAR][h]);
    __syncthreads();

    // windows around my points: rows y-4 .. y+4 at my columns, columns x-4 .. x+VX+3 of my row
    R yc[2 * RAD + 1][VX];
#pragma unroll
    for (int q = 0; q < 2 * RAD + 1; ++q)
      if (q != RAD) ld128s<R>(s + (warp + q) * SP + RAD + lane * VX, yc[q]);
    R xr[VX + 2 * RAD];
#pragma unroll
    for (int q = 0; q < (VX + 2 * RAD) / VX; ++q) {
      R t[VX];
      if (q * VX == RAD) {
#pragma unroll
        for (int e = 0; e < VX; ++e) t[e] = ring[(PH + 4) % RB][e];
      } else {
        ld128s<R>(s + (RAD + warp) * SP + lane * VX + q * VX, t);
      }
#pragma unroll
      for (int e = 0; e < VX; ++e) xr[q * VX + e] = t[e];
    }
    R cfr[NCA > 0 ? NCA : 1][VX];
    if constexpr (NCA > 0) {
      const R *cp = a.coef + off + (long long)z * g.pxy;
#pragma unroll
      for (int m = 0; m < NCA; ++m) {
#pragma unroll
        for (int e = 0; e < VX; ++e) cfr[m][e] = (R)0;
        if (ok) ld128<R>(cp + (long long)m * a.coef_stride, cfr[m]);
      }
    }
    R o[VX];
#pragma unroll
    for (int e = 0; e < VX; ++e) {
      RegNb4<R, PH> n{ring, xr, yc, e};
      if constexpr (NCA > 0) {
        RegCoef<R, NCA> cfp;
#pragma unroll
        for (int m = 0; m < NCA; ++m) cfp.v[m] = cfr[m][e];
        o[e] = StencilExpr<K>::template eval<R, FM>(n, cfp, (R)0, (R)0);
      } else {
        o[e] = StencilExpr<K>::template eval<R, FM>(n, a.cc, uo[PAR][e], rc[PAR][e]);
      }
    }
    R *outp = a.u + off + (long long)z * g.pxy;
    if (inter == (1u << VX) - 1u) {
      st128<R>(outp, o);
    } else if (inter != 0u) {
#pragma unroll
      for (int e = 0; e < VX; ++e)
        if ((inter >> e) & 1u) outp[e] = o[e];
    }
  };

  int z = zb;
  for (; z + RB <= ze; z += RB) {
    body(R4Phase<0>{}, z);     body(R4Phase<1>{}, z + 1); body(R4Phase<2>{}, z + 2);
    body(R4Phase<3>{}, z + 3); body(R4Phase<4>{}, z + 4); body(R4Phase<5>{}, z + 5);
    body(R4Phase<6>{}, z + 6); body(R4Phase<7>{}, z + 7); body(R4Phase<8>{}, z + 8);
    body(R4Phase<9>{}, z + 9);
  }
  if (z < ze) { body(R4Phase<0>{}, z); ++z; }
  if (z < ze) { body(R4Phase<1>{}, z); ++z; }
  if (z < ze) { body(R4Phase<2>{}, z); ++z; }
  if (z < ze) { body(R4Phase<3>{}, z); ++z; }
  if (z < ze) { body(R4Phase<4>{}, z); ++z; }
  if (z < ze) { body(R4Phase<5>{}, z); ++z; }
  if (z < ze) { body(R4Phase<6>{}, z); ++z; }
  if (z < ze) { body(R4Phase<7>{}, z); ++z; }
  if (z < ze) { body(R4Phase<8>{}, z); }
}


// accessor for the 9-plane ring of k_r4_async: ring[(PH + i) % 9] = plane z-4+i
template <typename R, int PH> struct RegNb4A {
  static constexpr int VX = Vec<R>::N;
  const R (*ring)[VX];
  const R *xr;
  const R (*yc)[VX];
  int e;
  template <int DX, int DY, int DZ> __device__ __forceinline__ R at() const {
    if constexpr (DZ != 0) return ring[(PH + 4 + DZ) % 9][e];
    else if constexpr (DX != 0) return xr[4 + e + DX];
    else if constexpr (DY != 0) return yc[4 + DY][e];
    else return ring[(PH + 4) % 9][e];
  }
};

// ------------------------------------------------------------------------------------------------------
// k_r4_async -- the same schedule with every HBM stream prefetched by cp.async (LDGSTS) instead of through
// registers.  The ring variant above keeps one plane of (v, u, roc2, halo) in flight per thread: 32 KB
// per SM, a third short of what 6.5 TB/s x ~1 us of loaded latency needs (Little's law), and deeper
// register prefetch costs the second CTA per SM.  Here NS = 4 shared-memory stages hold three planes in
// flight at no register cost:
//   * group G(p) = { halo strips of plane p -> staged plane buffer p % NS,
//                    u(p), roc2(p), v(p+4) of my own points -> private 16-byte slots of stage p % NS }
//     is issued three iterations before plane p is computed, one commit group per plane
//   * iteration z: cp.async.wait_group(NS-2) (G(z) has landed) -> my centre vector into buffer z % NS ->
//     __syncthreads -> issue G(z+3) (its buffer was read last in iteration z-1, which every warp has left)
//     -> compute plane z
// ------------------------------------------------------------------------------------------------------
#ifndef GIRIH_CUDA_EMU
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
#endif   // the test suite's CPU SIMT emulator (tests/cuda_emu) supplies its own queue-based versions

template <typename R, int NW> struct R4ACfg {
  using B = R4Cfg<R, NW>;
  static constexpr int NS = 4;                                  // stages: 3 planes in flight
  static constexpr int PLANE = B::SROWS * B::SP;                // elements per staged plane
  static constexpr int NPRIV = 3;                               // private streams: v(p+4), u(p), roc2(p)
  static constexpr size_t SMEM = ((size_t)NS * PLANE + (size_t)NS * NPRIV * B::NT * B::VX) * sizeof(R);
};

template <int K, typename R, int NW, bool FM = false>
__global__ void __launch_bounds__(32 * NW)
k_r4_async(const R4Args<R> a) {
  using Cfg = R4Cfg<R, NW>;
  using ACfg = R4ACfg<R, NW>;
  constexpr int RAD = Cfg::RAD, VX = Cfg::VX, WX = Cfg::WX, H = Cfg::H, NT = Cfg::NT;
  constexpr int SP = Cfg::SP, HXV = Cfg::HXV, NHV = Cfg::NHV, HPT = Cfg::HPT;
  constexpr int NS = ACfg::NS, PLANE = ACfg::PLANE, NPRIV = ACfg::NPRIV, RB = 9;
  constexpr int NCA = KTraits<K>::NCA;
  constexpr bool TO2 = KTraits<K>::TO == 2;
  static_assert(KTraits<K>::R == 4, "radius-4 operators only");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *planes = reinterpret_cast<R *>(smem_raw);                 // [NS][SROWS][SP]
  R *priv = planes + (size_t)NS * PLANE;                        // [NS][NPRIV][NT][VX]

  const DevGrid &g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0t = g.X0 + (int)blockIdx.x * WX, y0t = g.Y0 + (int)blockIdx.y * H;
  const int x = x0t + lane * VX, y = y0t + warp;
  const int zb = a.zb0 + (int)blockIdx.z * a.zchunk;
  const int ze = min(zb + a.zchunk, a.ze0);

  const bool ok = (x + VX <= g.px) && (y < g.ny_dev);
  unsigned inter = 0;
#pragma unroll
  for (int e = 0; e < VX; ++e)
    if ((x + e < g.X0 + g.nx) && (y < g.Y0 + g.ny)) inter |= 1u << e;
  const long long off = (long long)y * g.px + x;

  int hs_off[HPT];
  long long hg_off[HPT];
  bool h_ok[HPT];
#pragma unroll
  for (int h = 0; h < HPT; ++h) {
    const int item = tid + h * NT;
    int srow, scol;
    if (item < 2 * RAD * 32) {
      const int rr = item >> 5, vv = item & 31;
      srow = (rr < RAD) ? rr : H + rr;
      scol = RAD + vv * VX;
    } else {
      const int it2 = item - 2 * RAD * 32;
      const int row = it2 / (2 * HXV), k = it2 % (2 * HXV);
      const int side = k / HXV, hv = k % HXV;
      srow = RAD + row;
      scol = (side == 0) ? hv * VX : RAD + WX + hv * VX;
    }
    const int gx = x0t - RAD + scol, gy = y0t - RAD + srow;
    hs_off[h] = srow * SP + scol;
    hg_off[h] = (long long)gy * g.px + gx;
    h_ok[h] = (item < NHV) && (gx >= 0) && (gx + VX <= g.px) && (gy >= 0) && (gy < g.ny_dev);
  }

  auto priv_slot = [&](int stage, int which) -> R * {
    return priv + (((size_t)stage * NPRIV + which) * NT + tid) * VX;
  };
  // G(p): everything plane p needs from HBM / L2, as one commit group (possibly empty past the chunk)
  auto issue_group = [&](int p) {
    if (p < ze && p >= 0 && p < g.nz_dev) {
      const int st = (p - zb) & (NS - 1);
      R *buf = planes + (size_t)st * PLANE;
      const R *pv = a.v + (long long)p * g.pxy;
#pragma unroll
      for (int h = 0; h < HPT; ++h)
        if (h_ok[h]) cp_async16(buf + hs_off[h], pv + hg_off[h]);
      if (ok) {
        if (p + 4 < g.nz_dev) cp_async16(priv_slot(st, 0), a.v + off + (long long)(p + 4) * g.pxy);
        if constexpr (TO2) {
          cp_async16(priv_slot(st, 1), a.u + off + (long long)p * g.pxy);
          cp_async16(priv_slot(st, 2), a.roc2 + off + (long long)p * g.pxy);
        }
      }
    }
    cp_async_commit();
  };

  R ring[RB][VX];           // ring[(ph + i) % 9] = plane z-4+i (i = 0..7); (ph + 8) % 9 receives plane z+4
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int e = 0; e < VX; ++e) ring[i][e] = (R)0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (ok && zb - 4 + i >= 0 && zb - 4 + i < g.nz_dev) ld128<R>(a.v + off + (long long)(zb - 4 + i) * g.pxy, ring[i]);
  static_assert(NS == 4, "stage index uses a mask");
  issue_group(zb);
  issue_group(zb + 1);
  issue_group(zb + 2);

  auto body = [&](auto phase_tag, const int z) {
    constexpr int PH = decltype(phase_tag)::value;
    const int st = (z - zb) & (NS - 1);
    R *s = planes + (size_t)st * PLANE;
    cp_async_wait<NS - 2>();                       // G(z) complete (for this thread)
    // newest plane of my column, u(old) and roc2 of plane z from my private slots
    ld128s<R>(priv_slot(st, 0), ring[(PH + 8) % RB]);
    R uo[VX], rc[VX];
#pragma unroll
    for (int e = 0; e < VX; ++e) { uo[e] = (R)0; rc[e] = (R)0; }
    if constexpr (TO2) {
      ld128s<R>(priv_slot(st, 1), uo);
      ld128s<R>(priv_slot(st, 2), rc);
    }
    st128<R>(s + (RAD + warp) * SP + RAD + lane * VX, ring[(PH + 4) % RB]);
    __syncthreads();                               // halos of everyone + all centre vectors visible
    issue_group(z + NS - 1);                       // refills the buffer last read in iteration z-1

    R yc[2 * RAD + 1][VX];
#pragma unroll
    for (int q = 0; q < 2 * RAD + 1; ++q)
      if (q != RAD) ld128s<R>(s + (warp + q) * SP + RAD + lane * VX, yc[q]);
    R xr[VX + 2 * RAD];
#pragma unroll
    for (int q = 0; q < (VX + 2 * RAD) / VX; ++q) {
      R t[VX];
      if (q * VX == RAD) {
#pragma unroll
        for (int e = 0; e < VX; ++e) t[e] = ring[(PH + 4) % RB][e];
      } else {
        ld128s<R>(s + (RAD + warp) * SP + lane * VX + q * VX, t);
      }
#pragma unroll
      for (int e = 0; e < VX; ++e) xr[q * VX + e] = t[e];
    }
    R cfr[NCA > 0 ? NCA : 1][VX];
    if constexpr (NCA > 0) {
      const R *cp = a.coef + off + (long long)z * g.pxy;
#pragma unroll
      for (int m = 0; m < NCA; ++m) {
#pragma unroll
        for (int e = 0; e < VX; ++e) cfr[m][e] = (R)0;
        if (ok) ld128<R>(cp + (long long)m * a.coef_stride, cfr[m]);
      }
    }
    R o[VX];
#pragma unroll
    for (int e = 0; e < VX; ++e) {
      RegNb4A<R, PH> n{ring, xr, yc, e};
      if constexpr (NCA > 0) {
        RegCoef<R, NCA> cfp;
#pragma unroll
        for (int m = 0; m < NCA; ++m) cfp.v[m] = cfr[m][e];
        o[e] = StencilExpr<K>::template eval<R, FM>(n, cfp, (R)0, (R)0);
      } else {
        o[e] = StencilExpr<K>::template eval<R, FM>(n, a.cc, uo[e], rc[e]);
      }
    }
    R *outp = a.u + off + (long long)z * g.pxy;
    if (inter == (1u << VX) - 1u) {
      st128<R>(outp, o);
    } else if (inter != 0u) {
#pragma unroll
      for (int e = 0; e < VX; ++e)
        if ((inter >> e) & 1u) outp[e] = o[e];
    }
  };

  int z = zb;
  for (; z + RB <= ze; z += RB) {
    body(R4Phase<0>{}, z);     body(R4Phase<1>{}, z + 1); body(R4Phase<2>{}, z + 2);
    body(R4Phase<3>{}, z + 3); body(R4Phase<4>{}, z + 4); body(R4Phase<5>{}, z + 5);
    body(R4Phase<6>{}, z + 6); body(R4Phase<7>{}, z + 7); body(R4Phase<8>{}, z + 8);
  }
  if (z < ze) { body(R4Phase<0>{}, z); ++z; }
  if (z < ze) { body(R4Phase<1>{}, z); ++z; }
  if (z < ze) { body(R4Phase<2>{}, z); ++z; }
  if (z < ze) { body(R4Phase<3>{}, z); ++z; }
  if (z < ze) { body(R4Phase<4>{}, z); ++z; }
  if (z < ze) { body(R4Phase<5>{}, z); ++z; }
  if (z < ze) { body(R4Phase<6>{}, z); ++z; }
  if (z < ze) { body(R4Phase<7>{}, z); }
  cp_async_wait<0>();
}

}  // namespace girih
