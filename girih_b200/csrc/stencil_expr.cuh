// stencil_expr.cuh -- the per-point arithmetic of every operator, in ONE place.
//
// Each StencilExpr<K>::eval restates the FUNC_BODY() of GIRIH's src/kernels/stencils.c for table
// slot K with the reference's left-to-right association, using round-to-nearest add/mul
// intrinsics that the compiler never contracts into FMA (the library is also built with
// -fmad=false).  With that, every schedule in this library (naive, z-streamed, temporally fused,
// multi-GPU) evaluates the same expression on the same inputs as the reference's -O0 serial
// verifier (src/verification.c:315-479, 786-819) and the result is bit-identical.
//
// Accessor concepts
//   N  : template<int DX,int DY,int DZ> R at() const   -- neighbour of the point being updated
//   Cf : template<int M> R c() const                   -- coefficient m (scalar or per-point)
#pragma once
#include <cuda_runtime.h>

namespace girih {

template <typename R> struct Ar;
template <> struct Ar<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
};
template <> struct Ar<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
};

// Compile-time facts about each table slot (stencil_info_list[], src/kernels/stencils.c:260-271).
template <int K> struct KTraits;
template <> struct KTraits<0> { static constexpr int R = 4, TO = 2, NCA = 0,  NCS = 5; };
template <> struct KTraits<1> { static constexpr int R = 1, TO = 1, NCA = 0,  NCS = 2; };
template <> struct KTraits<2> { static constexpr int R = 1, TO = 1, NCA = 2,  NCS = 0; };
template <> struct KTraits<3> { static constexpr int R = 1, TO = 1, NCA = 4,  NCS = 0; };
template <> struct KTraits<4> { static constexpr int R = 4, TO = 1, NCA = 13, NCS = 0; };
template <> struct KTraits<5> { static constexpr int R = 1, TO = 1, NCA = 7,  NCS = 0; };
template <> struct KTraits<7> { static constexpr int R = 1, TO = 1, NCA = 0,  NCS = 4; };

#define NB(dx, dy, dz) (n.template at<dx, dy, dz>())
#define CF(m) (cf.template c<m>())
#define PSUM(ax, ay, az, bx, by, bz) A::add(NB(ax, ay, az), NB(bx, by, bz))

// Sum of products  c0*a0 + c1*a1 + c2*a2 + ...  evaluated left to right.
//   FM = false: every product and every sum rounded separately (reference built without FMA, and its
//               -O0 verifier) -- the default everywhere.
//   FM = true : the contraction gcc applies to the reference's FUNC_BODY under -O3 -mfma
//               (-ffp-contract=fast, gcc's default): its multiply-add pass visits the products in
//               statement order, so the FIRST product is fused onto the SECOND (which stays a plain
//               multiply) and every later product is fused onto the running sum:
//                   t = c1*a1;  t = fma(c0, a0, t);  t = fma(c2, a2, t);  ...
//               Pinned against oracle/_ref/ref_dump_*_fast (tests/test_oracle_vs_reference.py).
template <typename R, bool FM> struct Sop {
  using A = Ar<R>;
  static __device__ __forceinline__ R first(R c0, R a0, R c1, R a1) {
    if constexpr (FM) return A::fma(c0, a0, A::mul(c1, a1));
    else return A::add(A::mul(c0, a0), A::mul(c1, a1));
  }
  static __device__ __forceinline__ R next(R acc, R c, R a) {
    if constexpr (FM) return A::fma(c, a, acc);
    else return A::add(acc, A::mul(c, a));
  }
};

template <int K> struct StencilExpr;

// slot 1: 7-point constant coefficient -- src/kernels/stencils.c:73-78
template <> struct StencilExpr<1> {
  template <typename R, bool FM = false, typename N, typename Cf>
  static __device__ __forceinline__ R eval(const N &n, const Cf &cf, R, R) {
    using A = Ar<R>;
    using S = Sop<R, FM>;
    R acc = S::first(CF(0), NB(0, 0, 0), CF(1), PSUM(1, 0, 0, -1, 0, 0));
    acc = S::next(acc, CF(1), PSUM(0, 1, 0, 0, -1, 0));
    acc = S::next(acc, CF(1), PSUM(0, 0, -1, 0, 0, 1));
    return acc;
  }
};

// slot 2: 7-point variable coefficient (one off-centre coefficient) -- stencils.c:94-99
template <> struct StencilExpr<2> {
  template <typename R, bool FM = false, typename N, typename Cf>
  static __device__ __forceinline__ R eval(const N &n, const Cf &cf, R, R) {
    using A = Ar<R>;
    using S = Sop<R, FM>;
    const R c1 = CF(1);
    R acc = S::first(CF(0), NB(0, 0, 0), c1, PSUM(1, 0, 0, -1, 0, 0));
    acc = S::next(acc, c1, PSUM(0, 1, 0, 0, -1, 0));
    acc = S::next(acc, c1, PSUM(0, 0, 1, 0, 0, -1));
    return acc;
  }
};

// slot 3: 7-point variable, axis-symmetric coefficients -- stencils.c:121-126
template <> struct StencilExpr<3> {
  template <typename R, bool FM = false, typename N, typename Cf>
  static __device__ __forceinline__ R eval(const N &n, const Cf &cf, R, R) {
    using A = Ar<R>;
    using S = Sop<R, FM>;
    R acc = S::first(CF(0), NB(0, 0, 0), CF(1), PSUM(1, 0, 0, -1, 0, 0));
    acc = S::next(acc, CF(2), PSUM(0, 1, 0, 0, -1, 0));
    acc = S::next(acc, CF(3), PSUM(0, 0, 1, 0, 0, -1));
    return acc;
  }
};

// slot 5: 7-point variable, no symmetry -- stencils.c:193-201
template <> struct StencilExpr<5> {
  template <typename R, bool FM = false, typename N, typename Cf>
  static __device__ __forceinline__ R eval(const N &n, const Cf &cf, R, R) {
    using S = Sop<R, FM>;
    R acc = S::first(CF(0), NB(0, 0, 0), CF(1), NB(-1, 0, 0));
    acc = S::next(acc, CF(2), NB(1, 0, 0));
    acc = S::next(acc, CF(3), NB(0, -1, 0));
    acc = S::next(acc, CF(4), NB(0, 1, 0));
    acc = S::next(acc, CF(5), NB(0, 0, -1));
    acc = S::next(acc, CF(6), NB(0, 0, 1));
    return acc;
  }
};

// 25-point Laplacian-like sum shared by slots 0 and 4: distance m = 1..4, axes x, y, z in that
// order inside each m; coefficient index c0 + stride*(m-1) + axis*axis_stride.
// slot 0 (constant): coef[m] for all three axes  -> base 1, stride 1, axis_stride 0
// slot 4 (axsym):    COEF(1+3(m-1)+axis)         -> base 1, stride 3, axis_stride 1
#define STAR_YZ(acc, m, cy, cz)                            \
  acc = S::next(acc, CF(cy), PSUM(0, m, 0, 0, -m, 0));     \
  acc = S::next(acc, CF(cz), PSUM(0, 0, m, 0, 0, -m));
#define STAR_M(acc, m, cx, cy, cz)                         \
  acc = S::next(acc, CF(cx), PSUM(m, 0, 0, -m, 0, 0));     \
  STAR_YZ(acc, m, cy, cz)

// slot 0: 25-point constant, 2nd order in time -- stencils.c:27-41
//   u = 2*v - u + roc2*(lap); contracted form: fma(roc2, lap, fma(2, v, -u))
template <> struct StencilExpr<0> {
  template <typename R, bool FM = false, typename N, typename Cf>
  static __device__ __forceinline__ R eval(const N &n, const Cf &cf, R uold, R roc2) {
    using A = Ar<R>;
    using S = Sop<R, FM>;
    const R vc = NB(0, 0, 0);
    R lap = S::first(CF(0), vc, CF(1), PSUM(1, 0, 0, -1, 0, 0));
    STAR_YZ(lap, 1, 1, 1)
    STAR_M(lap, 2, 2, 2, 2)
    STAR_M(lap, 3, 3, 3, 3)
    STAR_M(lap, 4, 4, 4, 4)
    if constexpr (FM) return A::fma(roc2, lap, A::fma((R)2.0, vc, -uold));
    else return A::add(A::sub(A::mul((R)2.0, vc), uold), A::mul(roc2, lap));
  }
};

// slot 4: 25-point variable axis-symmetric, 1st order in time -- stencils.c:148-162
template <> struct StencilExpr<4> {
  template <typename R, bool FM = false, typename N, typename Cf>
  static __device__ __forceinline__ R eval(const N &n, const Cf &cf, R, R) {
    using A = Ar<R>;
    using S = Sop<R, FM>;
    R acc = S::first(CF(0), NB(0, 0, 0), CF(1), PSUM(1, 0, 0, -1, 0, 0));
    STAR_YZ(acc, 1, 2, 3)
    STAR_M(acc, 2, 4, 5, 6)
    STAR_M(acc, 3, 7, 8, 9)
    STAR_M(acc, 4, 10, 11, 12)
    return acc;
  }
};

// slot 7: 27-point box, constant -- stencils.c:227-242
template <> struct StencilExpr<7> {
  template <typename R, bool FM = false, typename N, typename Cf>
  static __device__ __forceinline__ R eval(const N &n, const Cf &cf, R, R) {
    using A = Ar<R>;
    using S = Sop<R, FM>;
    R acc = S::first(CF(0), NB(0, 0, 0), CF(1), PSUM(1, 0, 0, -1, 0, 0));
    acc = S::next(acc, CF(1), PSUM(0, 1, 0, 0, -1, 0));
    acc = S::next(acc, CF(1), PSUM(0, 0, -1, 0, 0, 1));
    acc = S::next(acc, CF(2), PSUM(1, 0, -1, -1, 0, -1));
    acc = S::next(acc, CF(2), PSUM(0, 1, -1, 0, -1, -1));
    acc = S::next(acc, CF(2), PSUM(1, 1, 0, -1, -1, 0));
    acc = S::next(acc, CF(2), PSUM(1, -1, 0, -1, 1, 0));
    acc = S::next(acc, CF(2), PSUM(1, 0, 1, -1, 0, 1));
    acc = S::next(acc, CF(2), PSUM(0, 1, 1, 0, -1, 1));
    acc = S::next(acc, CF(3), PSUM(1, 1, 1, -1, -1, -1));
    acc = S::next(acc, CF(3), PSUM(1, -1, 1, -1, 1, -1));
    acc = S::next(acc, CF(3), PSUM(-1, -1, 1, 1, 1, -1));
    acc = S::next(acc, CF(3), PSUM(-1, 1, 1, 1, -1, -1));
    return acc;
  }
};

#undef STAR_YZ
#undef STAR_M
#undef PSUM
#undef CF
#undef NB

}  // namespace girih
