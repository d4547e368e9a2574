// emu_runtime.cpp -- fiber scheduler of the CPU SIMT emulator (TEST INFRASTRUCTURE; see
// include/cuda_runtime.h).  One OS thread runs one CTA at a time; the CTA's CUDA threads are
// user-level fibers that run until their next synchronisation point (__syncthreads, warp
// shuffle/vote, cp.async wait) and then hand over to another fiber of the CTA.
//
// Scheduling order is a test parameter (CUDA_EMU_SCHED = 0 round robin, 1 reverse, 2 pseudo-random):
// a kernel whose result depends on the order in which warps or lanes reach a synchronisation point
// (a missing barrier, a shared-memory buffer reused too early) gives different bits under different
// orders, which the tests check for.
#include <cuda_runtime.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <sched.h>
#include <sys/mman.h>
#include <time.h>

#include <thread>
#include <vector>

namespace girih {
__thread __attribute__((aligned(128))) unsigned char smem_raw[cuda_emu::SMEM_BYTES];
}

namespace cuda_emu {

thread_local Block *B = nullptr;
thread_local Thread *TH = nullptr;
__thread int last_error = 0;
static int g_sched = 0;
static constexpr size_t STACK_BYTES = 256 * 1024;

#if !defined(__x86_64__)
#error "the fiber switch is written for x86-64 (System V ABI)"
#endif
// void emu_switch(void **save_sp, void *load_sp): saves the callee-saved registers of the running fiber
// on its stack, stores its stack pointer, loads the other fiber's stack pointer and registers.
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

static thread_local unsigned long long rng_state = 0x9e3779b97f4a7c15ull;
static inline unsigned rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (unsigned)(rng_state >> 32);
}

static void die(const char *what) {
  fprintf(stderr, "cuda_emu: %s (block %u,%u,%u)\n", what, B ? B->bid.x : 0, B ? B->bid.y : 0, B ? B->bid.z : 0);
  abort();
}

// pick the next runnable fiber (never the current one unless it is the only one left)
static int pick_next() {
  Block *b = B;
  const int n = b->nthreads;
  if (g_sched == 2) {
    for (int tries = 0; tries < 8; ++tries) {
      const int c = (int)(rnd() % (unsigned)n);
      if (!b->th[c].done && c != b->cur) return c;
    }
  }
  const int step = (g_sched == 1) ? n - 1 : 1;
  int c = b->cur;
  for (int i = 0; i < n; ++i) {
    c = (c + step) % n;
    if (!b->th[c].done) return c;
  }
  return -1;
}

static void switch_to(int next) {
  Block *b = B;
  Thread *from = TH;
  b->cur = next;
  TH = &b->th[next];
  emu_switch(&from->sp, TH->sp);
}

void yield() {
  Block *b = B;
  if (++b->idle > 64L * b->nthreads + 1024) die("deadlock: no fiber can make progress");
  const int next = pick_next();
  if (next < 0 || next == b->cur) return;
  switch_to(next);
}

static void barrier_release_if_complete(Block *b) {
  if (b->alive > 0 && b->bar_count >= b->alive) {
    b->bar_count = 0;
    b->bar_gen++;
  }
}

void syncthreads() {
  Block *b = B;
  const long gen = b->bar_gen;
  b->idle = 0;
  b->bar_count++;
  barrier_release_if_complete(b);
  while (b->bar_gen == gen) yield();
}

unsigned long long warp_exchange(unsigned long long v, int src_lane) {
  Block *b = B;
  Thread *t = TH;
  Warp &w = b->warps[t->warp];
  const unsigned n = t->ncoll++;
  const int slot = (int)(n & 1u);
  const long target = 32L * (long)(n / 2 + 1);
  w.buf[slot][t->lane] = v;
  w.arrived[slot]++;
  b->idle = 0;
  while (w.arrived[slot] < target) yield();
  return w.buf[slot][src_lane];
}

void cp_async_issue(void *dst, const void *src, int bytes) {
  Thread *t = TH;
  if (t->qn == t->qcap) {
    t->qcap = t->qcap ? 2 * t->qcap : 64;
    t->q = (Thread::Copy *)realloc(t->q, sizeof(Thread::Copy) * (size_t)t->qcap);
  }
  t->q[t->qn++] = Thread::Copy{dst, src, bytes};
}
void cp_async_commit_group() {
  Thread *t = TH;
  if (t->ngroups == 16) die("more than 16 cp.async groups in flight");
  t->gend[t->ngroups++] = t->qn;
}
// The copies of a group are performed as late as the programming model allows: when the thread waits
// for that group.  A kernel that reads a staged value too early sees stale shared memory here.
void cp_async_wait_group(int n) {
  Thread *t = TH;
  if (t->ngroups <= n) return;
  const int ndone = t->ngroups - n;
  const int upto = t->gend[ndone - 1];
  for (int i = 0; i < upto; ++i) memcpy(t->q[i].dst, t->q[i].src, (size_t)t->q[i].bytes);
  memmove(t->q, t->q + upto, sizeof(Thread::Copy) * (size_t)(t->qn - upto));
  t->qn -= upto;
  for (int g = ndone; g < t->ngroups; ++g) t->gend[g - ndone] = t->gend[g] - upto;
  t->ngroups -= ndone;
}

static void fiber_main() {
  Block *b = B;
  (*b->entry)();
  Thread *t = TH;
  if (t->ngroups != 0 || t->qn != 0) {
    // outstanding cp.async at thread exit complete implicitly on hardware; perform them
    t->gend[t->ngroups++] = t->qn;
    cp_async_wait_group(0);
  }
  t->done = true;
  b->alive--;
  b->idle = 0;
  barrier_release_if_complete(b);   // exited threads no longer take part in barriers
  const int next = pick_next();
  void *dummy;
  if (next < 0) emu_switch(&dummy, b->main_sp);
  else { b->cur = next; TH = &b->th[next]; emu_switch(&dummy, TH->sp); }
  die("a finished fiber was resumed");
}

static void prepare_fiber(Thread &t) {
  // initial frame: six callee-saved registers, then the return address emu_switch's `ret` jumps to
  uintptr_t top = ((uintptr_t)t.stack + STACK_BYTES) & ~(uintptr_t)15;
  void **a = (void **)(top - 16);    // 16-byte aligned slot: after `ret` rsp = top - 8, i.e. rsp % 16 == 8
  a[0] = (void *)&fiber_main;
  a[1] = nullptr;
  void **sp = a - 6;
  for (int i = 0; i < 6; ++i) sp[i] = nullptr;
  t.sp = (void *)sp;
}

struct Pool {   // per OS thread: fiber stacks and bookkeeping, reused from CTA to CTA
  Block blk{};
  int cap = 0;
  ~Pool() {
    for (int i = 0; i < cap; ++i) { munmap(blk.th[i].stack, STACK_BYTES); free(blk.th[i].q); }
    free(blk.th);
    free(blk.warps);
  }
  void reserve(int n) {
    if (n <= cap) return;
    blk.th = (Thread *)realloc(blk.th, sizeof(Thread) * (size_t)n);
    blk.warps = (Warp *)realloc(blk.warps, sizeof(Warp) * (size_t)((n + 31) / 32));
    for (int i = cap; i < n; ++i) {
      void *s = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (s == MAP_FAILED) { perror("cuda_emu: mmap"); abort(); }
      blk.th[i] = Thread{};
      blk.th[i].stack = (char *)s;
    }
    cap = n;
  }
};
static thread_local Pool pool;

static void run_block(uint3 bid, dim3 grid, dim3 block, const std::function<void()> &entry) {
  const int n = (int)(block.x * block.y * block.z);
  pool.reserve(n);
  Block *b = &pool.blk;
  b->nthreads = b->alive = n;
  b->bar_gen = 0;
  b->bar_count = 0;
  b->bid = bid;
  b->bdim = block;
  b->gdim = grid;
  b->idle = 0;
  b->entry = &entry;
  for (int w = 0; w < (n + 31) / 32; ++w) memset(&b->warps[w], 0, sizeof(Warp));
  for (int i = 0; i < n; ++i) {
    Thread &t = b->th[i];
    t.tid = uint3{(unsigned)i % block.x, ((unsigned)i / block.x) % block.y, (unsigned)i / (block.x * block.y)};
    t.lin = i; t.lane = i & 31; t.warp = i >> 5;
    t.done = false; t.ncoll = 0; t.qn = 0; t.ngroups = 0;
    prepare_fiber(t);
  }
  B = b;
  // shared memory is uninitialised on hardware: poison it so that a read before the first write shows
  memset(girih::smem_raw, 0xA5, sizeof(girih::smem_raw));
  const int first = (g_sched == 1) ? n - 1 : (g_sched == 2 ? (int)(rnd() % (unsigned)n) : 0);
  b->cur = first;
  TH = &b->th[first];
  emu_switch(&b->main_sp, TH->sp);
  if (b->alive != 0) die("CTA ended with live threads");
  TH = nullptr;
  B = nullptr;
}

void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()> &entry) {
  const long nthr = (long)block.x * block.y * block.z;
  if (nthr < 1 || nthr > 1024 || smem > SMEM_BYTES || grid.x < 1 || grid.y < 1 || grid.z < 1 || grid.y > 65535 ||
      grid.z > 65535) {
    last_error = cudaErrorInvalidValue;   // what cudaGetLastError() reports for an invalid configuration
    return;
  }
  // a partial last warp is accepted, but its threads must not take part in warp collectives (none of the
  // kernels launched that way -- one-thread flag kernels -- do)
  const char *e = getenv("CUDA_EMU_SCHED");
  g_sched = e ? atoi(e) : 0;
  const long nblocks = (long)grid.x * grid.y * grid.z;
#pragma omp parallel for schedule(dynamic, 1)
  for (long i = 0; i < nblocks; ++i) {
    rng_state = 0x9e3779b97f4a7c15ull ^ (unsigned long long)(i * 0x2545F4914F6CDD1Dull + 1);
    uint3 bid{(unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((long)grid.x * grid.y))};
    run_block(bid, grid, block, entry);
  }
}

// Cooperative launch: one OS thread per CTA, all running at once -- CTAs of kernels_r1x.cuh poll slots that other CTAs
// fill.  A fiber that spins on another CTA calls coop_pause(): the other fibers of its CTA get their turn, then the OS
// scheduler; the per-CTA deadlock counter does not apply (progress comes from outside), a wall-clock limit does.
static thread_local double coop_t0 = 0;
static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
void coop_pause() {
  Block *b = B;
  b->idle = 0;
  if (coop_t0 == 0) coop_t0 = now_s();
  else if (now_s() - coop_t0 > 300.0) die("cooperative launch: a CTA waited 300 s for another CTA");
  yield();
  sched_yield();
}

void launch_coop_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()> &entry) {
  const long nthr = (long)block.x * block.y * block.z;
  const long nblocks = (long)grid.x * grid.y * grid.z;
  if (nthr < 1 || nthr > 1024 || smem > SMEM_BYTES || nblocks < 1 || nblocks > 148) {
    last_error = nblocks > 148 ? cudaErrorCooperativeLaunchTooLarge : cudaErrorInvalidValue;
    return;
  }
  const char *e = getenv("CUDA_EMU_SCHED");
  g_sched = e ? atoi(e) : 0;
  std::vector<std::thread> th;
  th.reserve((size_t)nblocks);
  for (long i = 0; i < nblocks; ++i)
    th.emplace_back([=, &entry]() {
      rng_state = 0x9e3779b97f4a7c15ull ^ (unsigned long long)(i * 0x2545F4914F6CDD1Dull + 1);
      coop_t0 = 0;
      uint3 bid{(unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((long)grid.x * grid.y))};
      run_block(bid, grid, block, entry);
    });
  for (auto &t : th) t.join();
}

}  // namespace cuda_emu

unsigned __ballot_sync(unsigned, int pred) {
  using namespace cuda_emu;
  Block *b = B;
  Thread *t = TH;
  Warp &w = b->warps[t->warp];
  const unsigned n = t->ncoll++;
  const int slot = (int)(n & 1u);
  const long target = 32L * (long)(n / 2 + 1);
  w.buf[slot][t->lane] = pred ? 1ull : 0ull;
  w.arrived[slot]++;
  b->idle = 0;
  while (w.arrived[slot] < target) yield();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (unsigned)w.buf[slot][l] << l;
  return m;
}
