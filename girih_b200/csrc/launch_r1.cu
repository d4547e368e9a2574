// launch_r1.cu -- dispatch of the radius-1 launchers defined in inst_r1_k*_f*.cu
#include "launch.h"

namespace girih {
#define DECL(S) cudaError_t launch_r1_##S(int T, const StreamLaunch &s);
DECL(k1_f64) DECL(k1_f32) DECL(k2_f64) DECL(k2_f32) DECL(k3_f64) DECL(k3_f32) DECL(k5_f64) DECL(k5_f32)
#undef DECL

cudaError_t launch_r1(int kernel, int es, int T, const StreamLaunch &s) {
  switch (kernel) {
    case 1: return es == 8 ? launch_r1_k1_f64(T, s) : launch_r1_k1_f32(T, s);
    case 2: return es == 8 ? launch_r1_k2_f64(T, s) : launch_r1_k2_f32(T, s);
    case 3: return es == 8 ? launch_r1_k3_f64(T, s) : launch_r1_k3_f32(T, s);
    case 5: return es == 8 ? launch_r1_k5_f64(T, s) : launch_r1_k5_f32(T, s);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace girih
