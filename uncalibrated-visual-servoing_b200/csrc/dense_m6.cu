// dense_m6.cu — instantiates the dense (general covariance) kernels for M = 6 measurements.
#include "dense_kernels.cuh"
namespace uvs {
UVS_DEFINE_DENSE_LAUNCHERS(6)
}
