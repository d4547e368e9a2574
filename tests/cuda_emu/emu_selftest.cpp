// emu_selftest.cpp -- known-answer kernels for the emulator itself (TEST INFRASTRUCTURE): warp shuffles and
// votes, barrier semantics (a kernel WITHOUT its barrier must give order-dependent results, i.e. the
// emulator must be able to expose a missing barrier), cp.async group semantics, early thread exit.
#include <cuda_runtime.h>

namespace girih {
extern __thread __attribute__((aligned(128))) unsigned char smem_raw[];

__global__ void st_shuffle(int *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, v = 100 * warp + lane;
  const int up = __shfl_up_sync(0xffffffffu, v, 1), dn = __shfl_down_sync(0xffffffffu, v, 3);
  const double d = __shfl_down_sync(0xffffffffu, (double)v + 0.5, 1);
  const unsigned bal = __ballot_sync(0xffffffffu, lane % 3 == 0);
  const int any = __any_sync(0xffffffffu, lane == 7 && warp == 1), all = __all_sync(0xffffffffu, lane < 32);
  int *o = out + 8 * (blockIdx.x * blockDim.x + threadIdx.x);
  o[0] = up; o[1] = dn; o[2] = (int)(2 * d); o[3] = (int)bal; o[4] = any; o[5] = all;
  o[6] = __shfl_sync(0xffffffffu, v, 5); o[7] = (int)gridDim.x;
}

// every thread publishes a value, then reads its neighbour's from the other warp
template <bool BARRIER> __global__ void st_exchange(int *out, int rounds) {
  int *buf = reinterpret_cast<int *>(smem_raw);
  const int t = threadIdx.x, n = blockDim.x;
  int acc = 0;
  for (int r = 0; r < rounds; ++r) {
    buf[(r & 1) * n + t] = r * 1000 + t;
    if (BARRIER) __syncthreads();
    else (void)__shfl_sync(0xffffffffu, 0, 0);   // a yield point that is not a CTA barrier
    acc += buf[(r & 1) * n + (t + 37) % n];
  }
  out[blockIdx.x * n + t] = acc;
}

__global__ void st_cp_async(const double *src, int *out) {
  double *s = reinterpret_cast<double *>(smem_raw);
  const int t = threadIdx.x;
  s[2 * t] = -1.0; s[2 * t + 1] = -1.0;
  s[2 * (64 + t)] = -2.0; s[2 * (64 + t) + 1] = -2.0;
  cp_async16(s + 2 * t, src + 2 * t);
  cp_async_commit();
  cp_async16(s + 2 * (64 + t), src + 2 * (64 + t));
  cp_async_commit();
  const bool early_stale = (s[2 * t] == -1.0);   // nothing waited for: the emulator copies as late as allowed
  cp_async_wait<1>();
  const bool first_in = (s[2 * t] == src[2 * t]) && (s[2 * t + 1] == src[2 * t + 1]);
  const bool second_pending = (s[2 * (64 + t)] == -2.0);
  cp_async_wait<0>();
  const bool second_in = (s[2 * (64 + t)] == src[2 * (64 + t)]);
  __syncthreads();
  const bool neighbour = (s[2 * ((t + 1) % 64)] == src[2 * ((t + 1) % 64)]);
  out[t] = (early_stale ? 1 : 0) | (first_in ? 2 : 0) | (second_pending ? 4 : 0) | (second_in ? 8 : 0) |
           (neighbour ? 16 : 0);
}

__global__ void st_early_exit(int *out) {
  if (threadIdx.x >= 32) return;   // a whole warp leaves; the barrier below must not wait for it
  __syncthreads();
  out[threadIdx.x] = 1;
}
}  // namespace girih

extern "C" __attribute__((visibility("default"))) int emu_selftest(int which, int *out, const double *src,
                                                                   int rounds) {
  using namespace girih;
  switch (which) {
    case 0: { auto k = st_shuffle; GIRIH_LAUNCH(k, 2, 64, 0, nullptr, out); break; }
    case 1: { auto k = st_exchange<true>; GIRIH_LAUNCH(k, 3, 96, 2 * 96 * 4, nullptr, out, rounds); break; }
    case 2: { auto k = st_exchange<false>; GIRIH_LAUNCH(k, 3, 96, 2 * 96 * 4, nullptr, out, rounds); break; }
    case 3: { auto k = st_cp_async; GIRIH_LAUNCH(k, 1, 64, 4096, nullptr, src, out); break; }
    case 4: { auto k = st_early_exit; GIRIH_LAUNCH(k, 1, 64, 0, nullptr, out); break; }
    case 5: { auto k = st_early_exit; GIRIH_LAUNCH(k, 1, 2048, 0, nullptr, out); break; }   // invalid configuration
    default: return -1;
  }
  return cudaGetLastError();
}
