// dense_quad_m8.cu — instantiates the warp-specialised closed-loop kernel of the dense family (dense_quad.cuh) for M = 8.
#include "dense_quad.cuh"
namespace uvs {
UVS_DEFINE_QUAD_LAUNCHER(8)
}
