// explicit instantiation unit: radius-1 operator slot 2, float
#include "inst_r1.cuh"
namespace girih { GIRIH_INST_R1(2, float, k2_f32) }
