// twist_m2.cu — instantiates the camera-twist closed loop (tests/-script pipeline) for M = 2 measurements.
#include "twist_kernels.cuh"
namespace uvs {
UVS_DEFINE_TWIST_LAUNCHER(2)
}
