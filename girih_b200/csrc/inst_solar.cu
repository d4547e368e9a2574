// instantiation unit: the solar operator (slot 6), both precisions, both phases
#include <algorithm>

#include "kernels_solar.cuh"
#include "launch.h"

namespace girih {

template <typename R, int PHASE, bool HOIST, int SOLAR_BX, int SOLAR_BY>
static cudaError_t launch_solar_t(const SolarLaunch &s) {
  SolarArgs<R> a;
  a.u = (R *)s.u;
  a.coef = (const R *)s.coef;
  a.n2 = s.n2;
  a.nnx = s.nnx; a.nny = s.nny;
  a.xb = s.xb; a.xe = s.xe; a.yb = s.yb; a.ye = s.ye; a.zb = s.zb; a.ze = s.ze;
  const int nx = s.xe - s.xb, ny = s.ye - s.yb, nz = s.ze - s.zb;
  if (nx <= 0 || ny <= 0 || nz <= 0) return cudaSuccess;
  const long long tiles = (long long)((nx + SOLAR_BX - 1) / SOLAR_BX) * ((ny + SOLAR_BY - 1) / SOLAR_BY);
  // a chunk re-reads four source fields of one plane (the z carry): keep chunks long, but leave every SM several CTAs
  int zchunk = s.zchunk > 0 ? s.zchunk : 32;
  if (s.zchunk <= 0)
    while (zchunk > 4 && tiles * ((nz + zchunk - 1) / zchunk) < 148LL * 8 * 4) zchunk /= 2;
  zchunk = std::min(zchunk, nz);
  a.zchunk = zchunk;
  dim3 grid((nx + SOLAR_BX - 1) / SOLAR_BX, (ny + SOLAR_BY - 1) / SOLAR_BY, (nz + zchunk - 1) / zchunk);
  if (grid.y > 65535u || grid.z > 65535u) return cudaErrorInvalidValue;
  auto kfn = k_solar<R, PHASE, HOIST, SOLAR_BX, SOLAR_BY>;
  GIRIH_LAUNCH(kfn, grid, dim3(SOLAR_BX, SOLAR_BY, 1), 0, s.stream, a);
  return cudaGetLastError();
}

template <bool HOIST, int BX, int BY>
static cudaError_t launch_solar_v(int es, int phase, const SolarLaunch &s) {
  if (es == 8) return phase == 0 ? launch_solar_t<double, 0, HOIST, BX, BY>(s) : launch_solar_t<double, 1, HOIST, BX, BY>(s);
  return phase == 0 ? launch_solar_t<float, 0, HOIST, BX, BY>(s) : launch_solar_t<float, 1, HOIST, BX, BY>(s);
}

// tile option: 0 = default = 1; 1 = 64 x 4 CTA, own-cell loads component by component; 2 = 64 x 4 with all own-cell loads
// hoisted; 3 = 128 x 2, hoisted; 4 = 32 x 8, hoisted; 5 = 128 x 1, hoisted.  Measured on B200 (192^3, one time step,
// profiles/r02_solar_bench.log): fp64 1.120 / 1.447 / 1.715 / 1.453 / 1.300 ms, fp32 0.580 / 0.638 / 0.695 / 0.643 /
// 0.675 ms -- the plain schedule keeps 54-62 registers and twice the warps, and moves its two-phase traffic at
// 0.99 (fp64) / 0.96 (fp32) of the measured HBM copy bandwidth.
cudaError_t launch_solar(int es, int phase, const SolarLaunch &s) {
  switch (s.tile) {
    case 2: return launch_solar_v<true, 64, 4>(es, phase, s);
    case 3: return launch_solar_v<true, 128, 2>(es, phase, s);
    case 4: return launch_solar_v<true, 32, 8>(es, phase, s);
    case 5: return launch_solar_v<true, 128, 1>(es, phase, s);
    default: return launch_solar_v<false, 64, 4>(es, phase, s);
  }
}

}  // namespace girih
