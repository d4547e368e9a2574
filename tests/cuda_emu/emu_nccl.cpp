// emu_nccl.cpp -- runtime-API and NCCL stand-ins of the CPU emulator build (TEST INFRASTRUCTURE; see
// include/cuda_runtime.h).  Ranks are host threads of one process; a send is a buffered copy into the
// receiver's mailbox, a receive blocks until the matching send has been posted (per ordered pair, FIFO),
// which is the ordering NCCL guarantees for point-to-point calls on one communicator.
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <algorithm>
#include <mutex>
#include "nccl_dyn.h"

// ------------------------------------------------------------------------------------------------
// runtime API
// ------------------------------------------------------------------------------------------------
struct cuda_emu_event { std::chrono::steady_clock::time_point t; };

cudaError_t cudaGetDeviceCount(int *n) {
  const char *e = getenv("CUDA_EMU_DEVICES");
  *n = e ? atoi(e) : 8;
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t bytes) {
  *p = nullptr;
  if (posix_memalign(p, 256, bytes ? bytes : 256) != 0) return 2;   // cudaErrorMemoryAllocation
  memset(*p, 0xCD, bytes);   // device memory is not zeroed
  return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) {
  memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpy3D(const cudaMemcpy3DParms *p) {
  if (p->srcArray || p->dstArray) return cudaErrorInvalidValue;
  const cudaPitchedPtr &s = p->srcPtr, &d = p->dstPtr;
  if (p->srcPos.x + p->extent.width > s.pitch || p->dstPos.x + p->extent.width > d.pitch) return cudaErrorInvalidValue;
  if (p->srcPos.y + p->extent.height > s.ysize || p->dstPos.y + p->extent.height > d.ysize) return cudaErrorInvalidValue;
  for (size_t z = 0; z < p->extent.depth; ++z)
    for (size_t y = 0; y < p->extent.height; ++y) {
      const char *sp = (const char *)s.ptr + ((p->srcPos.z + z) * s.ysize + p->srcPos.y + y) * s.pitch + p->srcPos.x;
      char *dp = (char *)d.ptr + ((p->dstPos.z + z) * d.ysize + p->dstPos.y + y) * d.pitch + p->dstPos.x;
      memcpy(dp, sp, p->extent.width);
    }
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = malloc(1); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = malloc(1); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cuda_emu_event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
const char *cudaGetErrorString(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument";
    case 2: return "out of memory";
    default: return "emulated CUDA error";
  }
}

// ------------------------------------------------------------------------------------------------
// NCCL
// ------------------------------------------------------------------------------------------------
namespace {

struct World {
  int nranks = 0, joined = 0;
  std::mutex mu;
  std::condition_variable cv;
  std::map<std::pair<int, int>, std::deque<std::vector<char>>> box;   // (src, dst) -> messages in flight
  // all-reduce rendezvous
  int ar_count = 0;
  long ar_gen = 0;
  std::vector<long long> ar_vals;
  long long ar_result = 0;
};
std::mutex g_mu;
std::map<std::string, World *> g_worlds;
long g_next_id = 1;

struct Op { bool send; void *buf; size_t bytes; int peer; emu_nccl_comm *comm; };
thread_local int t_group_depth = 0;
thread_local std::vector<Op> t_ops;

size_t dt_size(ncclDataType_t t) {
  switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclFloat16: return 2;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    default: return 8;
  }
}

}  // namespace

struct emu_nccl_comm { World *w; int rank; };

static ncclResult_t run_ops(std::vector<Op> &ops) {
  // all sends first (buffered), then the receives in program order
  for (const Op &o : ops)
    if (o.send) {
      World *w = o.comm->w;
      std::lock_guard<std::mutex> lk(w->mu);
      w->box[{o.comm->rank, o.peer}].emplace_back((const char *)o.buf, (const char *)o.buf + o.bytes);
      w->cv.notify_all();
    }
  for (const Op &o : ops)
    if (!o.send) {
      World *w = o.comm->w;
      std::unique_lock<std::mutex> lk(w->mu);
      auto &q = w->box[{o.peer, o.comm->rank}];
      if (!w->cv.wait_for(lk, std::chrono::seconds(120), [&] { return !q.empty(); })) {
        fprintf(stderr, "emu_nccl: rank %d waited 120 s for a message from rank %d\n", o.comm->rank, o.peer);
        return ncclSystemError;
      }
      if (q.front().size() != o.bytes) {
        fprintf(stderr, "emu_nccl: rank %d expected %zu bytes from rank %d, got %zu\n", o.comm->rank, o.bytes, o.peer,
                q.front().size());
        return ncclInvalidArgument;
      }
      memcpy(o.buf, q.front().data(), o.bytes);
      q.pop_front();
    }
  ops.clear();
  return ncclSuccess;
}

static ncclResult_t e_GetUniqueId(ncclUniqueId *id) {
  std::lock_guard<std::mutex> lk(g_mu);
  memset(id, 0, sizeof(*id));
  snprintf(id->internal, sizeof(id->internal), "emu-world-%ld", g_next_id++);
  return ncclSuccess;
}
static ncclResult_t e_CommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
  World *w;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    World *&slot = g_worlds[std::string(id.internal)];
    if (!slot) { slot = new World(); slot->nranks = nranks; slot->ar_vals.assign((size_t)nranks, 0); }
    w = slot;
  }
  if (w->nranks != nranks || rank < 0 || rank >= nranks) return ncclInvalidArgument;
  *comm = new emu_nccl_comm{w, rank};
  std::unique_lock<std::mutex> lk(w->mu);
  w->joined++;
  w->cv.notify_all();
  if (!w->cv.wait_for(lk, std::chrono::seconds(120), [&] { return w->joined >= w->nranks; })) return ncclSystemError;
  return ncclSuccess;
}
static ncclResult_t e_CommDestroy(ncclComm_t c) { delete c; return ncclSuccess; }   // worlds are tiny and kept
static ncclResult_t e_GroupStart() { t_group_depth++; return ncclSuccess; }
static ncclResult_t e_GroupEnd() {
  if (t_group_depth <= 0) return ncclInvalidUsage;
  if (--t_group_depth == 0) return run_ops(t_ops);
  return ncclSuccess;
}
static ncclResult_t e_Send(const void *buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t c, cudaStream_t) {
  if (peer < 0 || peer >= c->w->nranks || peer == c->rank) return ncclInvalidArgument;
  t_ops.push_back(Op{true, const_cast<void *>(buf), count * dt_size(dt), peer, c});
  return t_group_depth > 0 ? ncclSuccess : run_ops(t_ops);
}
static ncclResult_t e_Recv(void *buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t c, cudaStream_t) {
  if (peer < 0 || peer >= c->w->nranks || peer == c->rank) return ncclInvalidArgument;
  t_ops.push_back(Op{false, buf, count * dt_size(dt), peer, c});
  return t_group_depth > 0 ? ncclSuccess : run_ops(t_ops);
}
static ncclResult_t e_AllReduce(const void *in, void *out, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t c,
                                cudaStream_t) {
  if (count != 1 || dt != ncclInt32) return ncclInvalidArgument;   // the one form the library uses
  World *w = c->w;
  std::unique_lock<std::mutex> lk(w->mu);
  const long gen = w->ar_gen;
  w->ar_vals[(size_t)c->rank] = *(const int *)in;
  if (++w->ar_count == w->nranks) {
    long long r = w->ar_vals[0];
    for (int i = 1; i < w->nranks; ++i) {
      const long long v = w->ar_vals[(size_t)i];
      r = (op == ncclMin) ? std::min(r, v) : (op == ncclMax) ? std::max(r, v) : (op == ncclSum) ? r + v : r * v;
    }
    w->ar_result = r;
    w->ar_count = 0;
    w->ar_gen++;
    w->cv.notify_all();
  } else if (!w->cv.wait_for(lk, std::chrono::seconds(120), [&] { return w->ar_gen != gen; })) {
    return ncclSystemError;
  }
  *(int *)out = (int)w->ar_result;
  return ncclSuccess;
}
static const char *e_GetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL error"; }

NcclDyn *emu_nccl_table() {
  static NcclDyn tab = {e_GetUniqueId, e_CommInitRank, e_CommDestroy, e_Send, e_Recv, e_AllReduce, e_GroupStart, e_GroupEnd,
                        e_GetErrorString};
  return &tab;
}
