#!/bin/bash
# ptxas register / spill report of the headline kernel only (dense_quad_kernel<8, GMCKF>: the kernel entry and its two
# non-inlined role functions): ~15 s instead of the full library build.
# usage: tools/spill_check.sh [extra nvcc flags, e.g. -DUVS_QUAD_MAXNREG=136 -DUVS_QUAD_GAIN_RB=3]
cd "$(dirname "$0")/../uncalibrated-visual-servoing_b200/csrc"
cat > /tmp/uvs_one.cu <<'EOT'
#include "dense_quad.cuh"
namespace uvs {
void* headline_ptr() { return (void*)dense_quad_kernel<8, UVS_METHOD_GMCKF>; }
}
EOT
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xptxas -v -I. "$@" \
    -c /tmp/uvs_one.cu -o /tmp/uvs_one.o 2>&1 | grep -E "Function properties|spill|Used" | grep -A1 -E "quad_(chain|covariance)_role|dense_quad_kernel|Used" \
    | sed -E 's/_ZN[0-9]+_INTERNAL_[a-z0-9_]+uvs[0-9]+//; s/ILi8ELi5.*//' | grep -v "^--"
