// blockdiag_m8.cu — instantiates the block-diagonal kernels for M = 8 measurements.
#include "blockdiag_kernels.cuh"
namespace uvs {
UVS_DEFINE_BLOCKDIAG_LAUNCHERS(8)
}
