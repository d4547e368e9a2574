// explicit instantiation unit: radius-1 operator slot 5, float
#include "inst_r1.cuh"
namespace girih { GIRIH_INST_R1(5, float, k5_f32) }
