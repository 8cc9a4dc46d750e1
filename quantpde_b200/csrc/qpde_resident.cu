// qpde_resident.cu — the RESIDENT family, common time loop: reverse time, uniform event times, no operator
// source (qpde_resident_impl.cuh holds the kernel; qpde_resident_gt.cu its general-time instantiations).
#define QPDE_RESIDENT_GT 0
#define QPDE_RESIDENT_LAUNCH launch_resident_common_time
#include "qpde_resident_impl.cuh"
