// cuda_runtime.h of the CPU SIMT emulator -- TEST INFRASTRUCTURE, never part of the product.
//
// The kernel sources under girih_b200/csrc are compiled a second time, unchanged, as plain C++ against
// this header (g++ -DGIRIH_CUDA_EMU -I tests/cuda_emu/include).  Every CUDA thread of a CTA becomes a
// user-level fiber; __syncthreads, warp shuffles/votes and cp.async are executed with their CUDA
// semantics by the small scheduler in emu_runtime.cpp.  This lets `pytest -m "not gpu"` run the real
// kernel code (tiling, masks, plane rotation, halo handling) against the oracle without a GPU; speed
// is irrelevant and nothing here is reachable from libgirih_cuda.so.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>

#ifndef GIRIH_CUDA_EMU
#error "this header is only for the emulator build of the test suite"
#endif

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ __thread   /* not thread_local: no dynamic-initialisation wrapper for extern declarations */
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorLaunchFailure = 719, cudaErrorCooperativeLaunchTooLarge = 720,
       cudaErrorNotSupported = 801 };
typedef void *cudaStream_t;
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaDevAttrMultiProcessorCount = 16 };

namespace cuda_emu {

constexpr size_t SMEM_BYTES = 227 * 1024;   // dynamic shared memory a CTA may ask for on sm_100

struct Thread {
  void *sp;          // saved stack pointer while the fiber is switched out
  char *stack;
  uint3 tid;
  int lin, lane, warp;
  bool done;
  unsigned ncoll;    // warp collectives executed so far
  // cp.async: copies of committed groups that have not been waited for yet
  struct Copy { void *dst; const void *src; int bytes; };
  Copy *q;
  int qcap, qn;            // copies queued (all groups)
  int gend[16], ngroups;   // end index of each committed group (FIFO)
};
struct Warp {
  unsigned long long buf[2][32];
  long arrived[2];
};
struct Block {
  Thread *th;
  Warp *warps;
  int nthreads, alive;
  long bar_gen;
  int bar_count;
  uint3 bid;
  dim3 bdim, gdim;
  void *main_sp;
  int cur;
  long idle;
  const std::function<void()> *entry;
};
extern thread_local Block *B;
extern thread_local Thread *TH;
extern __thread int last_error;

void yield();
void syncthreads();
unsigned long long warp_exchange(unsigned long long v, int src_lane);   // every lane deposits, reads src_lane
void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()> &entry);
// cooperative launch: every CTA of the grid runs at the same time (one OS thread per CTA), so CTAs may wait for each other
void launch_coop_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()> &entry);
void coop_pause();   // inside a spin on another CTA's progress: lets the other fibers of this CTA and the other CTAs run
void cp_async_issue(void *dst, const void *src, int bytes = 16);
void cp_async_commit_group();
void cp_async_wait_group(int n);

template <typename F, typename... A>
static inline void launch(F kfn, dim3 grid, dim3 block, size_t smem, cudaStream_t, A... args) {
  std::function<void()> entry = [=]() { kfn(args...); };
  launch_impl(grid, block, smem, entry);
}

template <typename F, typename A>
static inline cudaError_t launch_coop(F kfn, dim3 grid, dim3 block, size_t smem, cudaStream_t, A arg) {
  std::function<void()> entry = [=]() { kfn(arg); };
  launch_coop_impl(grid, block, smem, entry);
  const int e = last_error;
  last_error = 0;
  return e;
}

template <typename T> static inline unsigned long long to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  unsigned long long b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <typename T> static inline T from_bits(unsigned long long b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}

}  // namespace cuda_emu

#define GIRIH_LAUNCH(kfn, grid, block, smem, stream, ...) \
  cuda_emu::launch((kfn), dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__)
#define GIRIH_LAUNCH_COOP(kfn, grid, block, smem, stream, arg) \
  cuda_emu::launch_coop((kfn), dim3(grid), dim3(block), (size_t)(smem), (stream), (arg))

#define threadIdx (cuda_emu::TH->tid)
#define blockIdx (cuda_emu::B->bid)
#define blockDim (cuda_emu::B->bdim)
#define gridDim (cuda_emu::B->gdim)

static inline void __syncthreads() { cuda_emu::syncthreads(); }
// the lanes of a warp are fibers of one OS thread: a fiber switch orders their memory accesses.  One empty warp
// collective makes __syncwarp a real meeting point (a lane that runs ahead would otherwise read stale scratch).
static inline void __syncwarp() { (void)cuda_emu::warp_exchange(0ull, cuda_emu::TH->lane); }

// full-mask warp primitives (the only form the kernels use)
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  const int lane = cuda_emu::TH->lane, src = lane - (int)delta;
  const unsigned long long got = cuda_emu::warp_exchange(cuda_emu::to_bits(v), src < 0 ? lane : src);
  return cuda_emu::from_bits<T>(got);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  const int lane = cuda_emu::TH->lane, src = lane + (int)delta;
  const unsigned long long got = cuda_emu::warp_exchange(cuda_emu::to_bits(v), src > 31 ? lane : src);
  return cuda_emu::from_bits<T>(got);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
  return cuda_emu::from_bits<T>(cuda_emu::warp_exchange(cuda_emu::to_bits(v), src & 31));
}
unsigned __ballot_sync(unsigned mask, int pred);
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

template <typename T> static inline T __ldg(const T *p) { return *p; }

// IEEE round-to-nearest arithmetic, never contracted (the emulator is compiled with -ffp-contract=off)
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmaf_rn(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __fma_rn(double a, double b, double c) { return __builtin_fma(a, b, c); }

using std::max;
using std::min;

// the slice of the runtime API the launchers call
static inline cudaError_t cudaGetLastError() {
  const int e = cuda_emu::last_error;
  cuda_emu::last_error = 0;
  return e;
}
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int value) {
  return (size_t)value <= cuda_emu::SMEM_BYTES ? cudaSuccess : cudaErrorInvalidValue;
}
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int attr, int) {
  if (attr == cudaDevAttrMultiProcessorCount) *v = 148;
  return cudaSuccess;
}
template <typename F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *occ, F, int, size_t) {
  *occ = 1;
  return cudaSuccess;
}

// ---- the rest of the runtime API girih_cuda.cu uses: device memory is host memory, every stream
// ---- operation completes before the call returns (program order per rank), events carry wall-clock stamps
typedef struct cuda_emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct cudaPitchedPtr { void *ptr; size_t pitch, xsize, ysize; };
struct cudaPos { size_t x, y, z; };
struct cudaExtent { size_t width, height, depth; };
struct cudaMemcpy3DParms {
  void *srcArray; cudaPos srcPos; cudaPitchedPtr srcPtr;
  void *dstArray; cudaPos dstPos; cudaPitchedPtr dstPtr;
  cudaExtent extent; cudaMemcpyKind kind;
};
static inline cudaPitchedPtr make_cudaPitchedPtr(void *d, size_t p, size_t xsz, size_t ysz) { return cudaPitchedPtr{d, p, xsz, ysz}; }
static inline cudaPos make_cudaPos(size_t x, size_t y, size_t z) { return cudaPos{x, y, z}; }
static inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { return cudaExtent{w, h, d}; }

cudaError_t cudaGetDeviceCount(int *n);          // CUDA_EMU_DEVICES (default 8)
cudaError_t cudaSetDevice(int d);
cudaError_t cudaMalloc(void **p, size_t bytes);  // 256-byte aligned like the real allocator, poisoned
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
cudaError_t cudaFree(void *p);
enum { cudaHostAllocMapped = 2 };
static inline cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) { *p = calloc(1, bytes); return *p ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t bytes);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpy3D(const cudaMemcpy3DParms *p);
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t) { return cudaMemset(p, v, bytes); }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
static inline cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms *p, cudaStream_t) { return cudaMemcpy3D(p); }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t s);
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e);
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
const char *cudaGetErrorString(cudaError_t e);

// peer memory and IPC: every rank is a thread of this process, a handle is the pointer itself
enum { cudaErrorPeerAccessAlreadyEnabled = 704, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline long long clock64() { return (long long)__builtin_ia32_rdtsc(); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

// cp.async wrappers of kernels_r4.cuh and the split barrier of kernels_r1.cuh (the product versions are inline PTX)
namespace girih {
// LL slots of kernels_r1x.cuh: two 8-byte words {payload half, tag}, each written / read atomically
struct LLWord { unsigned lo, t0, hi, t1; };
static inline void ll_store(void *p, double v, unsigned tag) {
  unsigned long long b;
  memcpy(&b, &v, 8);
  unsigned long long *q = (unsigned long long *)p;
  __atomic_store_n(q, (b & 0xffffffffull) | ((unsigned long long)tag << 32), __ATOMIC_SEQ_CST);
  __atomic_store_n(q + 1, (b >> 32) | ((unsigned long long)tag << 32), __ATOMIC_SEQ_CST);
}
static inline LLWord ll_load(const void *p) {
  const unsigned long long *q = (const unsigned long long *)p;
  const unsigned long long w0 = __atomic_load_n(q, __ATOMIC_SEQ_CST), w1 = __atomic_load_n(q + 1, __ATOMIC_SEQ_CST);
  return LLWord{(unsigned)w0, (unsigned)(w0 >> 32), (unsigned)w1, (unsigned)(w1 >> 32)};
}
static inline double ll_value(const LLWord &w) {
  const unsigned long long b = (unsigned long long)w.lo | ((unsigned long long)w.hi << 32);
  double v;
  memcpy(&v, &b, 8);
  return v;
}
static inline LLWord ll_pack(double v, unsigned tag) {
  unsigned long long b;
  memcpy(&b, &v, 8);
  return LLWord{(unsigned)b, tag, (unsigned)(b >> 32), tag};
}
static inline void spin_pause() { cuda_emu::coop_pause(); }
template <int OFF> static inline void ll_store_if(void *p, double v, unsigned tag) { if (p) ll_store((unsigned char *)p + OFF, v, tag); }
static inline void ll_load_if(bool pred, const void *p, LLWord &w) { if (pred) w = ll_load(p); }
static inline void frame_load_if(bool pred, const double *p, double &v) { if (pred) v = *p; }
static inline void prefetch_l2_if(bool, const void *) {}
static inline void lds_if(bool pred, const double *p, double &v) { if (pred) v = *p; }
static inline void sts_if(bool pred, double *p, double v) { if (pred) *p = v; }
static inline int __double2loint(double v) { unsigned long long b; memcpy(&b, &v, 8); return (int)(unsigned)b; }
static inline int __double2hiint(double v) { unsigned long long b; memcpy(&b, &v, 8); return (int)(unsigned)(b >> 32); }
// mbarrier word: [phase:32][expected:16][pending:16]; all fibers of a CTA run on one OS thread
static inline void sb_init(unsigned long long *bar, int count) { *bar = ((unsigned long long)count << 16) | (unsigned)count; }
static inline void sb_arrive(unsigned long long *bar) {
  unsigned long long v = *bar;
  const unsigned expected = (unsigned)(v >> 16) & 0xffffu;
  unsigned pending = (unsigned)v & 0xffffu;
  unsigned long long phase = v >> 32;
  if (--pending == 0) { phase++; pending = expected; }
  *bar = (phase << 32) | ((unsigned long long)expected << 16) | pending;
  cuda_emu::B->idle = 0;
}
static inline void sb_wait(unsigned long long *bar, unsigned parity) {
  while ((unsigned)((*(volatile unsigned long long *)bar) >> 32 & 1u) == parity) cuda_emu::yield();
}
static inline void cp_async16(void *smem, const void *gmem) { cuda_emu::cp_async_issue(smem, gmem); }
template <int BYTES> static inline void cp_async_small(void *smem, const void *gmem) { cuda_emu::cp_async_issue(smem, gmem, BYTES); }
static inline void cp_async_commit() { cuda_emu::cp_async_commit_group(); }
template <int N> static inline void cp_async_wait() { cuda_emu::cp_async_wait_group(N); }
}  // namespace girih
