// explicit instantiation unit: radius-1 operator slot 1, double
#include "inst_r1.cuh"
namespace girih { GIRIH_INST_R1(1, double, k1_f64) }
