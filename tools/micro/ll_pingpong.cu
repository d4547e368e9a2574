// Micro-benchmark (measurement tool, not product): one-way latency of an LL slot hand-over between two CTAs on
// different SMs, as the exact-tiled sweep (kernels_r1x.cuh) uses it.  CTA 0 stores (value, tag) slots, CTA 1 polls
// until the tag shows up and answers; the round trip / 2 is printed in SM cycles for several partner CTAs.
#include <cstdio>
#include <cuda_runtime.h>
struct LLWord { unsigned lo, t0, hi, t1; };
__device__ __forceinline__ void ll_store(void *p, unsigned lo, unsigned hi, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag));
}
__device__ __forceinline__ LLWord ll_load(const void *p) {
  LLWord w;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.lo), "=r"(w.t0), "=r"(w.hi), "=r"(w.t1) : "l"(p));
  return w;
}
__global__ void pingpong(unsigned char *buf, int partner, int iters, long long *out) {
  // slot A: written by block 0, polled by block `partner`; slot B the other way.  Other blocks idle.
  unsigned char *A = buf, *B = buf + 4096;
  if (threadIdx.x != 0) return;
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
      ll_store(A, i, i, i);
      LLWord w = ll_load(B);
      while (w.t0 != (unsigned)i || w.t1 != (unsigned)i) w = ll_load(B);
    }
    out[0] = clock64() - t0;
  } else if ((int)blockIdx.x == partner) {
    for (int i = 1; i <= iters; ++i) {
      LLWord w = ll_load(A);
      while (w.t0 != (unsigned)i || w.t1 != (unsigned)i) w = ll_load(A);
      ll_store(B, i, i, i);
    }
  }
}
// plain L2 load latency (pointer chase with ld.relaxed.gpu) for comparison
__global__ void chase(unsigned long long *p, int iters, long long *out) {
  unsigned long long q = (unsigned long long)p;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) asm volatile("ld.relaxed.gpu.global.u64 %0, [%0];" : "+l"(q));
  out[0] = clock64() - t0;
  out[1] = (long long)q;
}
int main() {
  unsigned char *buf; long long *out, h[2];
  cudaMalloc(&buf, 1 << 16); cudaMemset(buf, 0, 1 << 16); cudaMalloc(&out, 16);
  const int iters = 2000;
  for (int partner : {1, 2, 17, 40, 74, 100, 147}) {
    cudaMemset(buf, 0, 1 << 16);
    pingpong<<<148, 32>>>(buf, partner, iters, out);
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("partner block %3d: one-way hand-over %.0f cycles (%s)\n", partner, (double)h[0] / iters / 2, cudaGetErrorString(cudaGetLastError()));
  }
  unsigned long long *p; cudaMalloc(&p, 8); unsigned long long self = (unsigned long long)p; cudaMemcpy(p, &self, 8, cudaMemcpyHostToDevice);
  chase<<<1, 1>>>(p, 2000, out); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("ld.relaxed.gpu pointer chase: %.0f cycles per load\n", (double)h[0] / 2000);
  return 0;
}
