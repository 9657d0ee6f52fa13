// blockdiag_m2.cu — instantiates the block-diagonal kernels for M = 2 measurements.
#include "blockdiag_kernels.cuh"
namespace uvs {
UVS_DEFINE_BLOCKDIAG_LAUNCHERS(2)
}
