// kernels_r4_strip.cuh -- strip variant of the z-streamed radius-4 sweep (PY rows per thread):
// the faster schedule for slot 4 (13 per-point coefficient arrays: bandwidth work with few
// arithmetic instructions per byte), where two rows per thread halve the shared-memory reads per
// point and the register column is rotated with moves.  kernels_r4.cuh holds the ring variant used
// for slot 0.
//
// Schedule (one CTA = NW warps, tile WX x (NW*PY), marching along z):
//   * a lane owns VX = 16 B / sizeof(Real) consecutive x points: 128-bit coalesced global access
//   * the z column (planes z-4 .. z+4 of the thread's own points) lives in registers and is
//     rotated once per plane; each plane of v is read from HBM exactly once per tile
//   * the centre plane z is staged in shared memory together with its 4-wide x/y halo strips
//     (double buffered, ONE __syncthreads per plane); x and y neighbours are then read back as
//     128-bit row/column windows that are shared by the VX x PY points of a thread
//   * u(old), roc2 and the per-point coefficients are touched at the thread's own points only,
//     so they stream straight from HBM into registers
// No temporal fusion here: with r = 4 the overlapped tile of a fused sweep wastes more than half
// of an SM-sized tile (halo 2*T*4 per axis) and slot 0 would need two extra arrays because its
// update reads the level it overwrites.  DESIGN.md, "r = 4 operators".
#pragma once
#include "common.cuh"
#include "stencil_expr.cuh"
#include "kernels_r4.cuh"   // R4Args

namespace girih {



template <typename R, int PY, int NW> struct R4StripCfg {
  static constexpr int RAD = 4;
  static constexpr int VX = Vec<R>::N;
  static constexpr int WX = 32 * VX;
  static constexpr int H = NW * PY;
  static constexpr int NT = 32 * NW;
  static constexpr int SP = WX + 2 * RAD;        // shared row pitch (elements), 16-byte multiple
  static constexpr int SROWS = H + 2 * RAD;
  static constexpr int HXV = RAD / VX;           // halo vectors per row side
  static constexpr int NHV = 2 * RAD * 32 + H * 2 * HXV;   // halo vectors per plane
  static constexpr int HPT = (NHV + NT - 1) / NT;           // halo vectors per thread
  static constexpr size_t SMEM = (size_t)2 * SROWS * SP * sizeof(R);
  static_assert(RAD % VX == 0, "halo must be whole vectors");
};

template <typename R, int PY> struct RegNb4Strip {
  static constexpr int VX = Vec<R>::N;
  const R (*zc)[PY][VX];     // zc[0..8] = planes z-4 .. z+4 at my points
  const R *xr;               // row window: x-4 .. x+VX+3 of row j
  const R (*yc)[VX];         // column window: rows y0-4 .. y0+PY+3
  int j, e;
  template <int DX, int DY, int DZ> __device__ __forceinline__ R at() const {
    if constexpr (DZ != 0) return zc[4 + DZ][j][e];
    else if constexpr (DX != 0) return xr[4 + e + DX];
    else if constexpr (DY != 0) return yc[4 + j + DY][e];
    else return zc[4][j][e];
  }
};

template <int K, typename R, int PY, int NW, bool FM = false>
__global__ void __launch_bounds__(32 * NW)
k_r4_strip(const R4Args<R> a) {
  using Cfg = R4StripCfg<R, PY, NW>;
  constexpr int RAD = Cfg::RAD, VX = Cfg::VX, WX = Cfg::WX, H = Cfg::H, NT = Cfg::NT;
  constexpr int SP = Cfg::SP, SROWS = Cfg::SROWS, HXV = Cfg::HXV, NHV = Cfg::NHV, HPT = Cfg::HPT;
  constexpr int NCA = KTraits<K>::NCA;
  static_assert(KTraits<K>::R == 4, "radius-4 operators only");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *sm = reinterpret_cast<R *>(smem_raw);   // [2][SROWS][SP]

  const DevGrid &g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0t = g.X0 + (int)blockIdx.x * WX, y0t = g.Y0 + (int)blockIdx.y * H;
  const int x = x0t + lane * VX, y0 = y0t + warp * PY;
  const int zb = a.zb0 + (int)blockIdx.z * a.zchunk;
  const int ze = min(zb + a.zchunk, a.ze0);

  const bool x_alloc = (x + VX <= g.px);
  bool row_alloc[PY];
  unsigned interior_xy = 0;
#pragma unroll
  for (int j = 0; j < PY; ++j) {
    row_alloc[j] = x_alloc && (y0 + j < g.ny_dev);
#pragma unroll
    for (int e = 0; e < VX; ++e)
      if ((x + e < g.X0 + g.nx) && (y0 + j < g.Y0 + g.ny)) interior_xy |= 1u << (j * VX + e);
  }
  const long long row0 = (long long)y0 * g.px + x;

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

  auto load_rows = [&](const R *base, int z, R (&dst)[PY][VX], bool coherent = false) {
    const bool zok = (z >= 0) && (z < g.nz_dev);
    const R *p = base + (long long)z * g.pxy + row0;
#pragma unroll
    for (int j = 0; j < PY; ++j) {
      if (zok && row_alloc[j]) {
        if (coherent) ld128g<R>(p + (long long)j * g.px, dst[j]);
        else ld128<R>(p + (long long)j * g.px, dst[j]);
      }
      else {
#pragma unroll
        for (int e = 0; e < VX; ++e) dst[j][e] = (R)0;
      }
    }
  };
  auto load_halo = [&](int z, R (&dst)[HPT][VX]) {
    const bool zok = (z >= 0) && (z < g.nz_dev);
    const R *p = a.v + (long long)z * g.pxy;
#pragma unroll
    for (int h = 0; h < HPT; ++h) {
      if (zok && h_ok[h]) ld128<R>(p + hg_off[h], dst[h]);
      else {
#pragma unroll
        for (int e = 0; e < VX; ++e) dst[h][e] = (R)0;
      }
    }
  };

  // z column: zc[i] = plane z-4+i.  Primed so that after the first rotation zc[0..7] = zb-4 .. zb+3
  R zc[9][PY][VX];
#pragma unroll
  for (int j = 0; j < PY; ++j)
#pragma unroll
    for (int e = 0; e < VX; ++e) zc[0][j][e] = (R)0;
#pragma unroll
  for (int i = 1; i < 9; ++i) load_rows(a.v, zb - 5 + i, zc[i]);
  R vnext[PY][VX], hnext[HPT][VX];
  load_rows(a.v, zb + 4, vnext);
  load_halo(zb, hnext);
  R uo_n[PY][VX], rc_n[PY][VX];
  if constexpr (KTraits<K>::TO == 2) {
    load_rows(a.u, zb, uo_n, true);
    load_rows(a.roc2, zb, rc_n);
  }

  for (int z = zb; z < ze; ++z) {
    const int cur = (z - zb) & 1;
    R *s = sm + (size_t)cur * SROWS * SP;

    // rotate the z column, take the prefetched plane z+4, start the next prefetches
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int e = 0; e < VX; ++e) zc[i][j][e] = zc[i + 1][j][e];
#pragma unroll
    for (int j = 0; j < PY; ++j)
#pragma unroll
      for (int e = 0; e < VX; ++e) zc[8][j][e] = vnext[j][e];
    R hcur[HPT][VX];
#pragma unroll
    for (int h = 0; h < HPT; ++h)
#pragma unroll
      for (int e = 0; e < VX; ++e) hcur[h][e] = hnext[h][e];
    R uo[PY][VX], rc[PY][VX];
    if constexpr (KTraits<K>::TO == 2) {
#pragma unroll
      for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int e = 0; e < VX; ++e) { uo[j][e] = uo_n[j][e]; rc[j][e] = rc_n[j][e]; }
    }
    if (z + 1 < ze) {
      load_rows(a.v, z + 5, vnext);
      load_halo(z + 1, hnext);
      if constexpr (KTraits<K>::TO == 2) {
        load_rows(a.u, z + 1, uo_n, true);
        load_rows(a.roc2, z + 1, rc_n);
      }
    }

    // stage plane z: my own points from the register column, my share of the halo strips
#pragma unroll
    for (int j = 0; j < PY; ++j)
      st128<R>(s + (RAD + warp * PY + j) * SP + RAD + lane * VX, zc[4][j]);
#pragma unroll
    for (int h = 0; h < HPT; ++h)
      if (tid + h * NT < NHV) st128<R>(s + hs_off[h], hcur[h]);
    __syncthreads();

    // column window shared by my PY rows: rows y0-4 .. y0+PY+3 at my VX columns
    R yc[PY + 2 * RAD][VX];
#pragma unroll
    for (int q = 0; q < PY + 2 * RAD; ++q) {
      if (q >= RAD && q < RAD + PY) {
#pragma unroll
        for (int e = 0; e < VX; ++e) yc[q][e] = zc[4][q - RAD][e];
      } else {
        ld128s<R>(s + (warp * PY + q) * SP + RAD + lane * VX, yc[q]);
      }
    }

    R *outp = a.u + (long long)z * g.pxy + row0;
#pragma unroll
    for (int j = 0; j < PY; ++j) {
      // row window x-4 .. x+VX+3
      R xr[VX + 2 * RAD];
#pragma unroll
      for (int q = 0; q < (VX + 2 * RAD) / VX; ++q) {
        R t[VX];
        ld128s<R>(s + (RAD + warp * PY + j) * SP + lane * VX + q * VX, t);
#pragma unroll
        for (int e = 0; e < VX; ++e) xr[q * VX + e] = t[e];
      }
      R cfr[NCA > 0 ? NCA : 1][VX];
      if constexpr (NCA > 0) {
        const R *cp = a.coef + (long long)z * g.pxy + row0 + (long long)j * g.px;
#pragma unroll
        for (int m = 0; m < NCA; ++m) {
          if (row_alloc[j]) ld128<R>(cp + (long long)m * a.coef_stride, cfr[m]);
          else {
#pragma unroll
            for (int e = 0; e < VX; ++e) cfr[m][e] = (R)0;
          }
        }
      }
      R o[VX];
#pragma unroll
      for (int e = 0; e < VX; ++e) {
        RegNb4Strip<R, PY> n{zc, xr, yc, j, e};
        if constexpr (NCA > 0) {
          RegCoef<R, NCA> cfp;
#pragma unroll
          for (int m = 0; m < NCA; ++m) cfp.v[m] = cfr[m][e];
          o[e] = StencilExpr<K>::template eval<R, FM>(n, cfp, (R)0, (R)0);
        } else {
          o[e] = StencilExpr<K>::template eval<R, FM>(n, a.cc, uo[j][e], rc[j][e]);
        }
      }
      const unsigned m = (interior_xy >> (j * VX)) & ((1u << VX) - 1u);
      if (m == (1u << VX) - 1u) {
        st128<R>(outp + (long long)j * g.px, o);
      } else if (m != 0u) {
#pragma unroll
        for (int e = 0; e < VX; ++e)
          if ((m >> e) & 1u) outp[(long long)j * g.px + e] = o[e];
      }
    }
  }
}

}  // namespace girih
