// explicit instantiation unit: radius-1 operator slot 5, double
#include "inst_r1.cuh"
namespace girih { GIRIH_INST_R1(5, double, k5_f64) }
