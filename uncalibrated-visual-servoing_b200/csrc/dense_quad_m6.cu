// dense_quad_m6.cu — instantiates the warp-specialised closed-loop kernel of the dense family (dense_quad.cuh) for M = 6.
#include "dense_quad.cuh"
namespace uvs {
UVS_DEFINE_QUAD_LAUNCHER(6)
}
