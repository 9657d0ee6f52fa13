#!/bin/bash
# ptxas register / spill report of the headline kernel only (dense_closed_loop_kernel<8, GMCKF, FAST>): ~40 s instead of
# the full library build.  usage: tools/spill_check.sh [extra nvcc flags, e.g. -DUVS_DENSE_MINB=3]
cd "$(dirname "$0")/../uncalibrated-visual-servoing_b200/csrc"
cat > /tmp/uvs_one.cu <<'EOT'
#include "dense_kernels.cuh"
namespace uvs {
void* headline_ptr() { return (void*)dense_closed_loop_kernel<8, UVS_METHOD_GMCKF, true>; }
}
EOT
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xptxas -v -I. "$@" \
    -c /tmp/uvs_one.cu -o /tmp/uvs_one.o 2>&1 | grep -E "spill|Used" | paste - -
