// aux_kernels.h — host-callable launchers of aux_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/uvs_b200.h"

namespace uvs {
cudaError_t launch_ibvs(int m, long long B, const double* J, const double* e, const double* kappa, double gain,
                        double* dq_out, int32_t* flags_out, cudaStream_t st);
cudaError_t launch_fk(long long B, const double* q, double* T06, double* Jg, cudaStream_t st);
cudaError_t launch_project(int nf, long long B, const double* q, const double* targets_host, double* f_out,
                           double* depth_out, double* X0_out, cudaStream_t st);
cudaError_t launch_noise_fill(const uvs_noise_params& np, unsigned long long trial0, long long B, int T, int m,
                              double* out, cudaStream_t st);
cudaError_t launch_start_poses(const double* q_start_host, unsigned long long pose_seed, int change,
                               unsigned long long trial0, long long B, double* q_out, cudaStream_t st);
cudaError_t launch_reduce_stats(long long B, unsigned long long trial0, int trial_stride, int sweep_epoch, int n_levels,
                                const double* stats, const int32_t* status, double* sums, double* n_fail,
                                cudaStream_t st);
cudaError_t run_fp64_peak(int iters, int sm_count, double* tflops, double* ms_out);
}  // namespace uvs
