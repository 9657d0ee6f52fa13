// twist_m8.cu — instantiates the camera-twist closed loop (tests/-script pipeline) for M = 8 measurements.
#include "twist_kernels.cuh"
namespace uvs {
UVS_DEFINE_TWIST_LAUNCHER(8)
}
