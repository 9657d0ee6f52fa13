// blockdiag_kernels.h — launchers of the block-diagonal kernel family (blockdiag_kernels.cu)
#pragma once
#include "kernel_args.cuh"

namespace uvs {
cudaError_t launch_blockdiag_replay(int m, const ReplayArgs& a, int sms, cudaStream_t st);
cudaError_t launch_blockdiag_closed_loop(int m, const ClosedLoopArgs& a, int sms, cudaStream_t st);
}  // namespace uvs
