// kernels_solar.cuh -- table slot 6, "solar": the 12-component complex field update of GIRIH
// (src/kernels/solar_spt_blk.ic:20-200 H update, :202-386 E update, :388-397 one time step = H then E).
//
// Data (the reference's own layout, uploaded unchanged): ONE array u of 12 fields x nnx*nny*nnz complex numbers
// (re, im interleaved; src/utils.c:168-172) and 28 complex coefficient arrays of the same shape (:199-201) -- 24 + 56
// reals per cell.  A time step reads and writes every field once and reads every coefficient once: 104 reals per cell
// (832 B in fp64), and with 13 arithmetic instructions per real moved the operator is HBM bound by a wide margin.
//
// Schedule: one kernel per phase (H, then E), because the phases are a true dependence over the whole grid (E at a cell
// reads the NEW H of its +x/+y/+z neighbours) and the update is in place.  Inside a phase every update reads only its
// own cell of the field it writes, plus two fields of the other kind that the phase does not touch, so the phase is a
// Jacobi sweep: any order gives the reference's bits.  A complex number is one 128-bit (fp64) / 64-bit (fp32) access;
// a thread owns one x position and marches along z -- upwards in the H phase, downwards in the E phase -- so the z
// neighbour of the source fields (z-1 for H, z+1 for E) is last iteration's own value and stays in registers; the x and y
// neighbours are the loads of the adjacent lane / the adjacent row of the same CTA and come from L1.
// HBM traffic per phase and cell: 6 fields in + 6 out + 6 source fields + 14 coefficients = 32 complex numbers; two
// phases 64 = 128 reals, i.e. 1.23 x the 104 of a (not in-place) fused step.
//
// Arithmetic: the reference's expressions with its left-to-right association, separately rounded (Ar<R>), see
// upd_h / upd_e below; bit-identical to the reference's production kernels and to its -O0 verifier
// (src/verification.c:481-784 holds the same expressions).
#pragma once
#include "common.cuh"
#include "stencil_expr.cuh"

namespace girih {

template <typename R> struct SolarArgs {
  R *u;                  // 12 fields, field f at u + f * n2
  const R *coef;         // 28 coefficient arrays, array m at coef + m * n2
  long long n2;          // 2 * nnx * nny * nnz
  int nnx, nny;          // row / plane pitch in cells
  int xb, xe, yb, ye, zb, ze;   // cells updated: [xb,xe) x [yb,ye) x [zb,ze)
  int zchunk;            // planes per CTA (blockIdx.z)
};

template <typename R> struct Cplx { R re, im; };

#ifdef GIRIH_CUDA_EMU
template <typename R> __device__ __forceinline__ Cplx<R> ldc(const R *p) { return Cplx<R>{p[0], p[1]}; }
template <typename R> __device__ __forceinline__ Cplx<R> ldw(const R *p) { return Cplx<R>{p[0], p[1]}; }
template <typename R> __device__ __forceinline__ void stc(R *p, Cplx<R> v) { p[0] = v.re; p[1] = v.im; }
#else
// read-only path: coefficients and the fields the running phase does not write
__device__ __forceinline__ Cplx<double> ldc(const double *p) {
  const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
  return Cplx<double>{t.x, t.y};
}
__device__ __forceinline__ Cplx<float> ldc(const float *p) {
  const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
  return Cplx<float>{t.x, t.y};
}
// the cell this thread is about to overwrite
__device__ __forceinline__ Cplx<double> ldw(const double *p) {
  const double2 t = *reinterpret_cast<const double2 *>(p);
  return Cplx<double>{t.x, t.y};
}
__device__ __forceinline__ Cplx<float> ldw(const float *p) {
  const float2 t = *reinterpret_cast<const float2 *>(p);
  return Cplx<float>{t.x, t.y};
}
__device__ __forceinline__ void stc(double *p, Cplx<double> v) { *reinterpret_cast<double2 *>(p) = make_double2(v.re, v.im); }
__device__ __forceinline__ void stc(float *p, Cplx<float> v) { *reinterpret_cast<float2 *>(p) = make_float2(v.re, v.im); }
#endif

// The four orders in which the reference writes the staggered difference of two source fields P, Q between the cell
// (i) and its neighbour (s):  A  P[i] - P[s] + Q[i] - Q[s]     B  P[s] - P[i] + Q[s] - Q[i]
//                             C  P[s] + Q[s] - P[i] - Q[i]     D  P[i] + Q[i] - P[s] - Q[s]
enum { SD_A = 0, SD_B = 1, SD_C = 2, SD_D = 3 };
template <int FORM, typename R> __device__ __forceinline__ R sdiff(R pi, R ps, R qi, R qs) {
  using A = Ar<R>;
  if constexpr (FORM == SD_A) return A::sub(A::add(A::sub(pi, ps), qi), qs);
  else if constexpr (FORM == SD_B) return A::sub(A::add(A::sub(ps, pi), qs), qi);
  else if constexpr (FORM == SD_C) return A::sub(A::sub(A::add(ps, qs), pi), qi);
  else return A::sub(A::sub(A::add(pi, qi), ps), qs);
}

// H component:  re = h.re*t.re - h.im*t.im [+ b.re] - c.re*dR + c.im*dI
//               im = h.re*t.im + h.im*t.re [+ b.im] - c.re*dI - c.im*dR      (solar_spt_blk.ic:82-83 and siblings)
template <bool BND, typename R>
__device__ __forceinline__ Cplx<R> upd_h(Cplx<R> h, Cplx<R> t, Cplx<R> c, Cplx<R> b, R dR, R dI) {
  using A = Ar<R>;
  R re = A::sub(A::mul(h.re, t.re), A::mul(h.im, t.im));
  R im = A::add(A::mul(h.re, t.im), A::mul(h.im, t.re));
  if constexpr (BND) {
    re = A::add(re, b.re);
    im = A::add(im, b.im);
  }
  re = A::add(A::sub(re, A::mul(c.re, dR)), A::mul(c.im, dI));
  im = A::sub(A::sub(im, A::mul(c.re, dI)), A::mul(c.im, dR));
  return Cplx<R>{re, im};
}
// E component:  re = e.re*t.re - e.im*t.im [+ b.re] + c.re*dR - c.im*dI
//               im = e.re*t.im + e.im*t.re [+ b.im] + c.re*dI + c.im*dR      (solar_spt_blk.ic:267-268 and siblings)
template <bool BND, typename R>
__device__ __forceinline__ Cplx<R> upd_e(Cplx<R> e, Cplx<R> t, Cplx<R> c, Cplx<R> b, R dR, R dI) {
  using A = Ar<R>;
  R re = A::sub(A::mul(e.re, t.re), A::mul(e.im, t.im));
  R im = A::add(A::mul(e.re, t.im), A::mul(e.im, t.re));
  if constexpr (BND) {
    re = A::add(re, b.re);
    im = A::add(im, b.im);
  }
  re = A::sub(A::add(re, A::mul(c.re, dR)), A::mul(c.im, dI));
  im = A::add(A::add(im, A::mul(c.re, dI)), A::mul(c.im, dR));
  return Cplx<R>{re, im};
}

// field numbers (solar_spt_blk.ic:29-41) and coefficient numbers (:44-60, :221-236)
enum { S_HYX = 0, S_HZX, S_HXY, S_HZY, S_HXZ, S_HYZ, S_EXZ, S_EYZ, S_EYX, S_EZX, S_EXY, S_EZY };
enum { S_HXBND = 12, S_HYBND = 13, S_EXBND = 26, S_EYBND = 27 };

// what one update reads at its own cell: the field it overwrites, its t and c coefficients, the boundary source
template <typename R> struct SolarIn { Cplx<R> f, t, c, b; };
template <int F, int CI, int TI, int BI, typename R>
__device__ __forceinline__ SolarIn<R> solar_load(const SolarArgs<R> &a, long long i) {
  SolarIn<R> in;
  in.f = ldw(a.u + (long long)F * a.n2 + i);
  in.t = ldc(a.coef + (long long)TI * a.n2 + i);
  in.c = ldc(a.coef + (long long)CI * a.n2 + i);
  in.b = Cplx<R>{R(0), R(0)};
  if constexpr (BI >= 0) in.b = ldc(a.coef + (long long)BI * a.n2 + i);
  return in;
}
// ... and the update itself: difference of the sources (pi, ps, qi, qs) in order FORM, new value stored in place
template <bool IS_H, int F, int BI, int FORM, typename R>
__device__ __forceinline__ void solar_finish(const SolarArgs<R> &a, long long i, const SolarIn<R> &in, Cplx<R> pi, Cplx<R> ps,
                                             Cplx<R> qi, Cplx<R> qs) {
  const R dR = sdiff<FORM>(pi.re, ps.re, qi.re, qs.re);
  const R dI = sdiff<FORM>(pi.im, ps.im, qi.im, qs.im);
  R *own = a.u + (long long)F * a.n2 + i;
  if constexpr (IS_H) stc(own, upd_h<(BI >= 0)>(in.f, in.t, in.c, in.b, dR, dI));
  else stc(own, upd_e<(BI >= 0)>(in.f, in.t, in.c, in.b, dR, dI));
}

// The six updates of a phase: X(n, field, c, t, boundary source or -1, order of the difference, P[i], P[s], Q[i], Q[s]).
//   H: sources (Exy, Exz) for Hy_x [z] and Hz_x [y]; (Eyx, Eyz) for Hx_y [z] and Hz_y [x]; (Ezx, Ezy) for Hx_z [y], Hy_z [x]
//   E: sources (Hzx, Hzy) for Ex_z [y] and Ey_z [x]; (Hxy, Hxz) for Ey_x [z] and Ez_x [y]; (Hyx, Hyz) for Ex_y [z], Ez_y [x]
#define SOLAR_H_LIST(X)                                                                      \
  X(0, S_HYX, 0, 6, S_HYBND, SD_A, exy, exy_z, exz, exz_z) /* solar_spt_blk.ic:78-84 */      \
  X(1, S_HZX, 1, 7, -1, SD_B, exy, exy_y, exz, exz_y)      /* :100-106 */                    \
  X(2, S_HXY, 2, 8, S_HXBND, SD_B, eyx, eyx_z, eyz, eyz_z) /* :122-128 */                    \
  X(3, S_HZY, 3, 9, -1, SD_A, eyx, eyx_x, eyz, eyz_x)      /* :144-150 */                    \
  X(4, S_HXZ, 4, 10, -1, SD_A, ezx, ezx_y, ezy, ezy_y)     /* :166-172 */                    \
  X(5, S_HYZ, 5, 11, -1, SD_C, ezx, ezx_x, ezy, ezy_x)     /* :188-194 */
#define SOLAR_E_LIST(X)                                                                      \
  X(0, S_EXZ, 14, 20, -1, SD_B, hzx, hzx_y, hzy, hzy_y)      /* solar_spt_blk.ic:263-269 */  \
  X(1, S_EYZ, 15, 21, -1, SD_D, hzx, hzx_x, hzy, hzy_x)      /* :285-291 */                  \
  X(2, S_EYX, 16, 22, S_EYBND, SD_B, hxy, hxy_z, hxz, hxz_z) /* :307-313 */                  \
  X(3, S_EZX, 17, 23, -1, SD_D, hxy, hxy_y, hxz, hxz_y)      /* :329-335 */                  \
  X(4, S_EXY, 18, 24, S_EXBND, SD_A, hyx, hyx_z, hyz, hyz_z) /* :351-357 */                  \
  X(5, S_EZY, 19, 25, -1, SD_B, hyx, hyx_x, hyz, hyz_x)      /* :373-379 */
#define SOLAR_LOAD(n, F, CI, TI, BI, FORM, pi, ps, qi, qs) const SolarIn<R> in##n = solar_load<F, CI, TI, BI>(a, i);
#define SOLAR_FIN_H(n, F, CI, TI, BI, FORM, pi, ps, qi, qs) solar_finish<true, F, BI, FORM>(a, i, in##n, pi, ps, qi, qs);
#define SOLAR_FIN_E(n, F, CI, TI, BI, FORM, pi, ps, qi, qs) solar_finish<false, F, BI, FORM>(a, i, in##n, pi, ps, qi, qs);
#define SOLAR_ONE_H(n, F, CI, TI, BI, FORM, pi, ps, qi, qs) \
  { SOLAR_LOAD(n, F, CI, TI, BI, FORM, pi, ps, qi, qs) SOLAR_FIN_H(n, F, CI, TI, BI, FORM, pi, ps, qi, qs) }
#define SOLAR_ONE_E(n, F, CI, TI, BI, FORM, pi, ps, qi, qs) \
  { SOLAR_LOAD(n, F, CI, TI, BI, FORM, pi, ps, qi, qs) SOLAR_FIN_E(n, F, CI, TI, BI, FORM, pi, ps, qi, qs) }

// PHASE 0: H update (neighbours at -x / -y / -z, marching z upwards); PHASE 1: E update (+x / +y / +z, downwards).
// HOIST: all 20 own-cell loads of an iteration are issued before the first update (34 independent 128-bit loads in
// flight per thread with the 14 source loads) instead of component by component -- slower on B200 (86-108 registers,
// half the warps; inst_solar.cu has the numbers), kept as a tuner candidate; BX x BY = x positions x rows per CTA.
template <typename R, int PHASE, bool HOIST, int BX, int BY>
__global__ void __launch_bounds__(BX *BY) k_solar(const SolarArgs<R> a) {
  const int x = a.xb + (int)blockIdx.x * BX + (int)threadIdx.x;
  const int y = a.yb + (int)blockIdx.y * BY + (int)threadIdx.y;
  if (x >= a.xe || y >= a.ye) return;   // no barrier below
  const int z0 = a.zb + (int)blockIdx.z * a.zchunk;
  const int z1 = (z0 + a.zchunk < a.ze) ? z0 + a.zchunk : a.ze;
  if (z0 >= z1) return;
  constexpr int D = PHASE == 0 ? -1 : 1;            // where the neighbour sits
  const long long SX = 2LL * D, SY = 2LL * D * a.nnx, SZ = 2LL * D * a.nnx * a.nny;
  auto src = [&](int f, long long i) { return ldc(a.u + (long long)f * a.n2 + i); };
  int z = PHASE == 0 ? z0 : z1 - 1;
  long long i = 2LL * (((long long)z * a.nny + y) * a.nnx + x);
  if constexpr (PHASE == 0) {
    Cplx<R> exy_z = src(S_EXY, i + SZ), exz_z = src(S_EXZ, i + SZ), eyx_z = src(S_EYX, i + SZ), eyz_z = src(S_EYZ, i + SZ);
    for (; z < z1; ++z, i -= SZ) {
      const Cplx<R> exy = src(S_EXY, i), exz = src(S_EXZ, i), eyx = src(S_EYX, i), eyz = src(S_EYZ, i);
      const Cplx<R> ezx = src(S_EZX, i), ezy = src(S_EZY, i);
      const Cplx<R> exy_y = src(S_EXY, i + SY), exz_y = src(S_EXZ, i + SY), ezx_y = src(S_EZX, i + SY), ezy_y = src(S_EZY, i + SY);
      const Cplx<R> eyx_x = src(S_EYX, i + SX), eyz_x = src(S_EYZ, i + SX), ezx_x = src(S_EZX, i + SX), ezy_x = src(S_EZY, i + SX);
      if constexpr (HOIST) {
        SOLAR_H_LIST(SOLAR_LOAD)
        SOLAR_H_LIST(SOLAR_FIN_H)
      } else {
        SOLAR_H_LIST(SOLAR_ONE_H)
      }
      exy_z = exy; exz_z = exz; eyx_z = eyx; eyz_z = eyz;
    }
  } else {
    Cplx<R> hxy_z = src(S_HXY, i + SZ), hxz_z = src(S_HXZ, i + SZ), hyx_z = src(S_HYX, i + SZ), hyz_z = src(S_HYZ, i + SZ);
    for (; z >= z0; --z, i -= SZ) {
      const Cplx<R> hzx = src(S_HZX, i), hzy = src(S_HZY, i), hxy = src(S_HXY, i), hxz = src(S_HXZ, i);
      const Cplx<R> hyx = src(S_HYX, i), hyz = src(S_HYZ, i);
      const Cplx<R> hzx_y = src(S_HZX, i + SY), hzy_y = src(S_HZY, i + SY), hxy_y = src(S_HXY, i + SY), hxz_y = src(S_HXZ, i + SY);
      const Cplx<R> hzx_x = src(S_HZX, i + SX), hzy_x = src(S_HZY, i + SX), hyx_x = src(S_HYX, i + SX), hyz_x = src(S_HYZ, i + SX);
      if constexpr (HOIST) {
        SOLAR_E_LIST(SOLAR_LOAD)
        SOLAR_E_LIST(SOLAR_FIN_E)
      } else {
        SOLAR_E_LIST(SOLAR_ONE_E)
      }
      hxy_z = hxy; hxz_z = hxz; hyx_z = hyx; hyz_z = hyz;
    }
  }
}

}  // namespace girih
