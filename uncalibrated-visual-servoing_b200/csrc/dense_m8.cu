// dense_m8.cu — instantiates the dense (general covariance) kernels for M = 8 measurements.
#include "dense_kernels.cuh"
namespace uvs {
UVS_DEFINE_DENSE_LAUNCHERS(8)
}
