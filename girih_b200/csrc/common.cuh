// common.cuh -- device-side layout descriptor and small helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace girih {

// Layout of every domain-sized array in HBM (one z-slab of one GPU).
//
//   element (x, y, z) of the device array lives at  (z * ny_dev + y) * px + x
//
// The device layout is NOT the host layout: the interior origin (X0, Y0, Z0) is placed so that
// x = X0 is 128-byte aligned, rows are padded to a multiple of 128 bytes, and guard rows/planes
// surround the reference's r-deep Dirichlet frame so that overlapped tiles and deep (T*r) halos
// can be read without bounds checks on the hot path.  Host index (i, j, k) of the reference layout
// (src/kernels/stencils.h:29-32) maps to device (i - r + X0, j - r + Y0, k - r + Z0).
struct DevGrid {
  int px;          // row pitch in elements (multiple of 128 B / sizeof(Real))
  int ny_dev;      // rows per plane
  int nz_dev;      // planes
  long long pxy;   // plane pitch in elements
  int X0, Y0, Z0;  // device coordinates of the first interior point
  int nx, ny, nz;  // LOCAL interior extent
  int r;           // stencil radius
  // Global-interior range in z for the pass-through mask: a point is updated only if it is an
  // interior point of the GLOBAL domain.  Slabs that have a neighbour recompute the neighbour's
  // planes inside their deep halo, so the bound on that side is open (+-2^30).
  int zlo, zhi;
};

// Scalar coefficients of the constant-coefficient operators, passed by value (constant bank).
template <typename R> struct ConstCoef {
  R v[5];
  template <int M> __device__ __forceinline__ R c() const { return v[M]; }
};

// Per-point coefficients: array m lives `stride` elements after array m-1 (same DevGrid layout).
template <typename R> struct PointCoef {
  const R *__restrict__ p;   // element (x,y,z) of coefficient array 0
  long long stride;
  template <int M> __device__ __forceinline__ R c() const { return __ldg(p + (long long)M * stride); }
};

// Per-point coefficients already sitting in registers.
template <typename R, int NCA> struct RegCoef {
  R v[NCA > 0 ? NCA : 1];
  template <int M> __device__ __forceinline__ R c() const { return v[M]; }
};

// Neighbour accessor straight from global memory (naive kernel, box step).
template <typename R> struct GlobalNb {
  const R *__restrict__ p;   // centre element
  int px;
  long long pxy;
  template <int DX, int DY, int DZ> __device__ __forceinline__ R at() const {
    return __ldg(p + DX + (long long)DY * px + (long long)DZ * pxy);
  }
};

template <typename R> struct Vec;   // 128-bit vector of R
template <> struct Vec<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

// 128-bit streaming load / store (read-only path, no L1 allocation for write-once data)
template <typename R> __device__ __forceinline__ void ld128(const R *p, R (&v)[Vec<R>::N]);
template <> __device__ __forceinline__ void ld128<double>(const double *p, double (&v)[2]) {
  const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
  v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld128<float>(const float *p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
// 128-bit coherent global load, for arrays the same kernel also writes (slot 0 reads u(t-1) and
// overwrites it with u(t+1) at the same address)
template <typename R> __device__ __forceinline__ void ld128g(const R *p, R (&v)[Vec<R>::N]);
template <> __device__ __forceinline__ void ld128g<double>(const double *p, double (&v)[2]) {
  const double2 t = *reinterpret_cast<const double2 *>(p);
  v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld128g<float>(const float *p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4 *>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
// 128-bit load from shared memory (plain ld, not the read-only path)
template <typename R> __device__ __forceinline__ void ld128s(const R *p, R (&v)[Vec<R>::N]);
template <> __device__ __forceinline__ void ld128s<double>(const double *p, double (&v)[2]) {
  const double2 t = *reinterpret_cast<const double2 *>(p);
  v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld128s<float>(const float *p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4 *>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <typename R> __device__ __forceinline__ void st128(R *p, const R (&v)[Vec<R>::N]);
template <> __device__ __forceinline__ void st128<double>(double *p, const double (&v)[2]) {
  *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
}
template <> __device__ __forceinline__ void st128<float>(float *p, const float (&v)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

}  // namespace girih
