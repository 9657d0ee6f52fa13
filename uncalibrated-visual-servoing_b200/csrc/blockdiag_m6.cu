// blockdiag_m6.cu — instantiates the block-diagonal kernels for M = 6 measurements.
#include "blockdiag_kernels.cuh"
namespace uvs {
UVS_DEFINE_BLOCKDIAG_LAUNCHERS(6)
}
