// explicit instantiation unit: radius-1 operator slot 1, float
#include "inst_r1.cuh"
namespace girih { GIRIH_INST_R1(1, float, k1_f32) }
