// launch.h -- interface between the context (girih_cuda.cu) and the per-operator kernel
// translation units (inst_*.cu), which are compiled separately so the build parallelises.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

// Every kernel of the library is started through this one macro.  The CPU SIMT emulator of the test
// suite (tests/cuda_emu, test infrastructure only) compiles the same kernel and launcher sources with
// its own definition; the product build always takes the <<<>>> form below.
#ifndef GIRIH_LAUNCH
#define GIRIH_LAUNCH(kfn, grid, block, smem, stream, ...) (kfn)<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// Cooperative launch (all CTAs of the grid co-resident): the exact-tiled sweep (kernels_r1x.cuh), whose CTAs wait for
// each other's rim values.  `arg` is the kernel's single by-value argument struct.
#ifndef GIRIH_LAUNCH_COOP
#define GIRIH_LAUNCH_COOP(kfn, grid, block, smem, stream, arg) \
  girih::launch_cooperative((const void *)(kfn), (grid), (block), (smem), (stream), (void *)&(arg))
namespace girih {
static inline cudaError_t launch_cooperative(const void *f, dim3 grid, dim3 block, size_t smem, cudaStream_t s, void *arg) {
  void *args[1] = {arg};
  return cudaLaunchCooperativeKernel(f, grid, block, args, smem, s);
}
}  // namespace girih
#endif

namespace girih {

struct StreamLaunch {
  DevGrid g;
  const void *in;        // array holding the newest level (read)
  void *out;             // array written (slot 0: also read at the centre, level before `in`)
  const void *roc2;      // slot 0
  const void *coef;      // per-point coefficient arrays
  long long coef_stride;
  double cc[5];          // scalar coefficients
  int zb0, ze0;          // output planes [zb0, ze0), device z
  int zb1, ze1;          // optional second range swept by the same launch (fused r = 1 sweep only), else 0, 0
  int zchunk;            // 0 = choose
  int tile;              // 0 = default, else PY*100 + NW
  int variant;           // 0 = auto, 2 = force the fused-sweep kernel also for T = 1
  int contract;          // 1 = FMA-contracted arithmetic (stencil_expr.cuh, Sop<R, true>); default tiles only
  // halo push (fused r = 1 sweep of slot 1 only): neighbours' output arrays in peer memory, pre-shifted (see R1Args);
  // nullptr / empty plane ranges when the pass pushes nothing
  void *push_up = nullptr, *push_dn = nullptr;
  int push_up_from = 0x7fffffff, push_dn_below = -0x7fffffff;
  // exact-tiled sweep (kernels_r1x.cuh): inbound edge slots of the context, the running slot tag, the give-up flag, and
  // how many CTAs the device keeps resident at once.  nullptr = the context has none: overlapped tiles only
  unsigned char *xbuf = nullptr;
  size_t xbuf_bytes = 0;
  unsigned *xseq = nullptr;
  int *xerr = nullptr;
  int nsm = 148;
  cudaStream_t stream;
};

// is `tile` one of the exact-tiled (non-overlapping, edge hand-off) variants?  10000 + PY*100 + NW
static inline bool tile_is_exact(int tile) { return tile >= 10000 && tile < 20000; }

// T fused steps of a radius-1 operator (slots 1, 2, 3, 5); es = sizeof(real)
cudaError_t launch_r1(int kernel, int es, int T, const StreamLaunch &a);
// one step of a radius-4 operator (slots 0, 4)
cudaError_t launch_r4(int kernel, int es, const StreamLaunch &a);
// one step of the 27-point box operator (slot 7)
cudaError_t launch_box(int es, const StreamLaunch &a);

// one phase (0 = H update, 1 = E update; a time step is both, in that order) of the solar operator (slot 6) over a box
// of cells of the reference's own array layout (12 complex fields in one array, 28 complex coefficient arrays)
struct SolarLaunch {
  void *u;
  const void *coef;
  long long n2;                   // 2 * nnx * nny * nnz: reals per field / coefficient array
  int nnx, nny;
  int xb, xe, yb, ye, zb, ze;     // cells [xb,xe) x [yb,ye) x [zb,ze), host-array coordinates
  int zchunk;                     // 0 = choose
  int tile;                       // 0 = default, 1..5 = schedule variants (inst_solar.cu)
  cudaStream_t stream;
};
cudaError_t launch_solar(int es, int phase, const SolarLaunch &a);

}  // namespace girih
