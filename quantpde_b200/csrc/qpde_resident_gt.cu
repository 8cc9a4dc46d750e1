// qpde_resident_gt.cu — the RESIDENT family with the general time loop: forward-time sweeps
// (TimeIteration<true>), explicit event times, an operator source b(t) (examples/experimental/glwb.cpp).
#define QPDE_RESIDENT_GT 1
#define QPDE_RESIDENT_LAUNCH launch_resident_general_time
#include "qpde_resident_impl.cuh"
