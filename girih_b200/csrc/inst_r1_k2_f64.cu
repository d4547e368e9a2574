// explicit instantiation unit: radius-1 operator slot 2, double
#include "inst_r1.cuh"
namespace girih { GIRIH_INST_R1(2, double, k2_f64) }
