// placeholder until the RESIDENT family lands
#include "qpde_kernels.cuh"
namespace qpde {
bool resident_supported(const DeviceBatch &) { return false; }
int launch_resident(const DeviceBatch &, const DeviceResult &, cudaStream_t, int, long long *) { return QPDE_ERR_UNSUPPORTED; }
}
