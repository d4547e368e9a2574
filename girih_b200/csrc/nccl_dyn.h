// nccl_dyn.h -- NCCL resolved at run time with dlopen("libnccl.so.2").
//
// The library must load on machines without NCCL (and on the CPU-only build container, where
// the non-GPU tests check that every symbol of include/girih_cuda.h is exported), so it carries
// no DT_NEEDED entry for NCCL.  Inside a PyTorch process the already-loaded bundled libnccl is
// picked up by its SONAME; the stand-alone mwd_kernel binary finds the system copy.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types and enums only

struct NcclDyn {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  const char *(*GetErrorString)(ncclResult_t);
};

static char g_nccl_err[256] = "not attempted";
static inline const char *nccl_dyn_error() { return g_nccl_err; }

#ifdef GIRIH_CUDA_EMU
// test suite's CPU emulator build (tests/cuda_emu): ranks are host threads, NCCL is an in-process mailbox
NcclDyn *emu_nccl_table();
static inline NcclDyn *nccl_dyn() { return emu_nccl_table(); }
#else
static inline NcclDyn *nccl_dyn() {
  static NcclDyn tab;
  static int state = 0;   // 0 = not tried, 1 = ok, -1 = failed
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (state == 0) {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      snprintf(g_nccl_err, sizeof(g_nccl_err), "%s", dlerror());
      state = -1;
    } else {
      bool ok = true;
#define SYM(field, name)                                                   \
  do {                                                                     \
    *(void **)(&tab.field) = dlsym(h, name);                               \
    if (!tab.field) { ok = false; snprintf(g_nccl_err, sizeof(g_nccl_err), "missing symbol %s", name); } \
  } while (0)
      SYM(GetUniqueId, "ncclGetUniqueId");
      SYM(CommInitRank, "ncclCommInitRank");
      SYM(CommDestroy, "ncclCommDestroy");
      SYM(Send, "ncclSend");
      SYM(Recv, "ncclRecv");
      SYM(AllReduce, "ncclAllReduce");
      SYM(GroupStart, "ncclGroupStart");
      SYM(GroupEnd, "ncclGroupEnd");
      SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
      state = ok ? 1 : -1;
    }
  }
  return state == 1 ? &tab : nullptr;
}
#endif
