// twist_m6.cu — instantiates the camera-twist closed loop (tests/-script pipeline) for M = 6 measurements.
#include "twist_kernels.cuh"
namespace uvs {
UVS_DEFINE_TWIST_LAUNCHER(6)
}
