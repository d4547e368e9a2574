// explicit instantiation unit: radius-1 operator slot 3, float
#include "inst_r1.cuh"
namespace girih { GIRIH_INST_R1(3, float, k3_f32) }
