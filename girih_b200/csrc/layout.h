// layout.h -- host-side computation of the DevGrid layout of one z-slab (see common.cuh).
// One function, used by girih_gpu_create() and by the test suite's CPU SIMT emulator, so that both lay
// arrays out identically.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace girih {

// Fills `g` for a slab of st[0] x st[1] x st[2] interior points of a radius-r operator whose steppers
// fuse at most max_tfuse steps; returns the number of guard/halo planes kept on each z side.
//   * interior x origin 128-byte aligned, rows padded to 128 bytes plus one spare 128-byte group so that
//     the last (partial) tile's vector accesses stay inside the row
//   * z-slab runs keep up to 4 fused passes' worth of halo planes so that one exchange can serve several
//     passes (run_passes); a single slab needs only the pipeline's own guard planes
static inline int make_dev_grid(DevGrid &g, const int st[3], int r, int max_tfuse, int elem_size, int rank,
                                int nranks) {
  const int epl = 128 / elem_size;            // elements per 128 bytes
  const int guard = std::max(max_tfuse * r, r);   // deepest halo / overlap any stepper uses
  g.r = r; g.nx = st[0]; g.ny = st[1]; g.nz = st[2];
  g.X0 = epl;                                 // >= guard, keeps x = X0 line aligned
  g.Y0 = guard;
  const int zguard = guard * (nranks > 1 ? 4 : 1);
  g.Z0 = zguard;
  g.px = (g.X0 + g.nx + guard + epl - 1) / epl * epl + epl;
  g.ny_dev = g.Y0 + g.ny + guard;
  g.nz_dev = g.Z0 + g.nz + zguard;
  g.pxy = (long long)g.px * g.ny_dev;
  g.zlo = (rank == 0) ? g.Z0 : -(1 << 30);
  g.zhi = (rank == nranks - 1) ? g.Z0 + g.nz : (1 << 30);
  return zguard;
}

}  // namespace girih
