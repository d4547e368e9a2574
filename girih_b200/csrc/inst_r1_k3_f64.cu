// explicit instantiation unit: radius-1 operator slot 3, double
#include "inst_r1.cuh"
namespace girih { GIRIH_INST_R1(3, double, k3_f64) }
