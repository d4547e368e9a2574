// kernels_naive.cuh -- one thread per lattice site, every neighbour straight from global memory.
//
// This is the box-step operator behind girih_gpu_step_box (the spt_blk_func_t contract,
// src/kernels/stencils_spt_blk.ic:19-48) and the "variant 1" stepper kept for cross-checking the
// streamed kernels on the device.  It is not the performance path.
#pragma once
#include "common.cuh"
#include "stencil_expr.cuh"

namespace girih {

template <int K, typename R, bool FM = false>
__global__ void __launch_bounds__(256)
k_naive(DevGrid g, R *__restrict__ u, const R *__restrict__ v, const R *__restrict__ roc2,
        const R *__restrict__ coef, long long coef_stride, ConstCoef<R> cc,
        int xb, int yb, int zb, int xe, int ye, int ze) {
  const int x = xb + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = yb + blockIdx.y * blockDim.y + threadIdx.y;
  const int z = zb + blockIdx.z;
  if (x >= xe || y >= ye || z >= ze) return;
  const long long idx = ((long long)z * g.ny_dev + y) * g.px + x;
  GlobalNb<R> n{v + idx, g.px, g.pxy};
  R uold = (R)0, rc = (R)0;
  if constexpr (KTraits<K>::TO == 2) { uold = u[idx]; rc = __ldg(roc2 + idx); }
  R out;
  if constexpr (KTraits<K>::NCA > 0) {
    PointCoef<R> cf{coef + idx, coef_stride};
    out = StencilExpr<K>::template eval<R, FM>(n, cf, uold, rc);
  } else {
    out = StencilExpr<K>::template eval<R, FM>(n, cc, uold, rc);
  }
  u[idx] = out;
}

}  // namespace girih
