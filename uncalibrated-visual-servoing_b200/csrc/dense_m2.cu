// dense_m2.cu — instantiates the dense (general covariance) kernels for M = 2 measurements.
#include "dense_kernels.cuh"
namespace uvs {
UVS_DEFINE_DENSE_LAUNCHERS(2)
}
