// blockdiag_kernels.cu — dispatch of the block-diagonal kernel family by measurement count (kernels: blockdiag_kernels.cuh,
// one translation unit per M: blockdiag_m{2,6,8}.cu).
#include "blockdiag_kernels.h"

namespace uvs {

template <int M> cudaError_t bd_replay_m(const ReplayArgs& a, int sms, cudaStream_t st);
template <int M> cudaError_t bd_closed_m(const ClosedLoopArgs& a, int sms, cudaStream_t st);
extern template cudaError_t bd_replay_m<2>(const ReplayArgs&, int, cudaStream_t);
extern template cudaError_t bd_replay_m<6>(const ReplayArgs&, int, cudaStream_t);
extern template cudaError_t bd_replay_m<8>(const ReplayArgs&, int, cudaStream_t);
extern template cudaError_t bd_closed_m<2>(const ClosedLoopArgs&, int, cudaStream_t);
extern template cudaError_t bd_closed_m<6>(const ClosedLoopArgs&, int, cudaStream_t);
extern template cudaError_t bd_closed_m<8>(const ClosedLoopArgs&, int, cudaStream_t);

cudaError_t launch_blockdiag_replay(int m, const ReplayArgs& a, int sms, cudaStream_t st) {
    switch (m) {
    case 2: return bd_replay_m<2>(a, sms, st);
    case 6: return bd_replay_m<6>(a, sms, st);
    case 8: return bd_replay_m<8>(a, sms, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_blockdiag_closed_loop(int m, const ClosedLoopArgs& a, int sms, cudaStream_t st) {
    switch (m) {
    case 2: return bd_closed_m<2>(a, sms, st);
    case 6: return bd_closed_m<6>(a, sms, st);
    case 8: return bd_closed_m<8>(a, sms, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace uvs
