// aux_kernels.cu — small stand-alone kernels behind the C ABI: IBVS, UR10 kinematics, projection,
// initial guess, noise streams, start poses, statistics reduction and the FP64 peak probe.
#include "common.cuh"
#include "philox.cuh"
#include "smallmat.cuh"
#include "ur10.cuh"
#include "aux_kernels.h"

namespace uvs {

// ---- IBVS: one problem per warp (experiment.py:300-312) ------------------------------------------
template <int M>
__global__ void ibvs_kernel(long long B, const double* J, const double* e, const double* kappa, double gain,
                            double* dq_out, int32_t* flags_out) {
    __shared__ __align__(16) double sJ[8][M * 6];
    __shared__ __align__(16) double sc[8][56];
    __shared__ __align__(16) double sBig[8][(M >= 6) ? pinv_scratch_doubles(M) : 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long b = (long long)blockIdx.x * 8 + warp; b < B; b += (long long)gridDim.x * 8) {
        for (int i = lane; i < M * 6; i += 32) sJ[warp][i] = J[b * M * 6 + i];
        double bl = 0.0;
        if (lane < M) bl = (kappa ? kappa[b * M + lane] : 1.0) * e[b * M + lane];
        __syncwarp();
        double x[6];
        int flags = 0;
        pinv_warp<M>(sJ[warp], bl, sc[warp], lane, x, flags, sBig[warp]);
#pragma unroll
        for (int r = 0; r < 6; ++r) if (lane == r) dq_out[b * 6 + r] = -gain * x[r];
        if (flags_out && lane == 0) flags_out[b] = flags;
        __syncwarp();
    }
}

cudaError_t launch_ibvs(int m, long long B, const double* J, const double* e, const double* kappa, double gain,
                        double* dq_out, int32_t* flags_out, cudaStream_t st) {
    const int grid = (int)((B + 7) / 8 < 4096 ? (B + 7) / 8 : 4096);
    switch (m) {
    case 2: ibvs_kernel<2><<<grid, 256, 0, st>>>(B, J, e, kappa, gain, dq_out, flags_out); break;
    case 6: ibvs_kernel<6><<<grid, 256, 0, st>>>(B, J, e, kappa, gain, dq_out, flags_out); break;
    case 8: ibvs_kernel<8><<<grid, 256, 0, st>>>(B, J, e, kappa, gain, dq_out, flags_out); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ---- UR10 forward kinematics + geometric Jacobian (ur10_simulation.py:97-139) ---------------------
__global__ void fk_kernel(long long B, const double* q, double* T06, double* Jg) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double qq[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) qq[j] = q[b * 6 + j];
    UR10 kin;
    Tf T6;
    ur10_fk(qq, T6, &kin);
    if (T06) {
        double* o = T06 + b * 16;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) o[i * 4 + j] = T6.r[i][j];
            o[i * 4 + 3] = T6.p[i];
        }
        o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 1.0;
    }
    if (Jg) {
        double J[6][6];
        ur10_jacobian(kin, J);
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) Jg[b * 36 + i * 6 + j] = J[i][j];
    }
}

cudaError_t launch_fk(long long B, const double* q, double* T06, double* Jg, cudaStream_t st) {
    fk_kernel<<<(int)((B + 127) / 128), 128, 0, st>>>(B, q, T06, Jg);
    return cudaGetLastError();
}

struct Targets { double p[12]; };

// ---- projection / initial guess (Appendix A pixel model; experiment.py:94-114) --------------------
__global__ void project_kernel(int nf, long long B, const double* q, Targets tg, double* f_out, double* depth_out,
                               double* X0_out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double qq[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) qq[j] = q[b * 6 + j];
    UR10 kin;
    Tf T6;
    ur10_fk(qq, T6, &kin);
    double Jg[6][6];
    if (X0_out) ur10_jacobian(kin, Jg);
    const int m = 2 * nf;
    for (int i = 0; i < nf; ++i) {
        double z;
        const double u = project_coord(T6, tg.p + 3 * i, 0, z);
        const double v = project_coord(T6, tg.p + 3 * i, 1, z);
        if (f_out) { f_out[b * m + 2 * i] = u; f_out[b * m + 2 * i + 1] = v; }
        if (depth_out) depth_out[b * nf + i] = z;
        if (X0_out) {
            const double zc = camera_distance(T6, tg.p + 3 * i);
            for (int ax = 0; ax < 2; ++ax) {
                double row[6];
                initial_guess_row(kin, Jg, 2 * i + ax, u, v, zc, row);
#pragma unroll
                for (int j = 0; j < 6; ++j) X0_out[b * 6 * m + (2 * i + ax) * 6 + j] = row[j];
            }
        }
    }
}

cudaError_t launch_project(int nf, long long B, const double* q, const double* targets_host, double* f_out,
                           double* depth_out, double* X0_out, cudaStream_t st) {
    Targets tg;
    for (int i = 0; i < 12; ++i) tg.p[i] = i < 3 * nf ? targets_host[i] : 0.0;
    project_kernel<<<(int)((B + 63) / 64), 64, 0, st>>>(nf, B, q, tg, f_out, depth_out, X0_out);
    return cudaGetLastError();
}

// ---- noise streams (noise.py:81-207): one thread per (trial, feature pair) -------------------------
__global__ void noise_fill_kernel(uvs_noise_params np, unsigned long long trial0, long long B, int T, int m,
                                  double* out) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs = m / 2;
    if (id >= B * pairs) return;
    const long long b = id / pairs;
    const int pr = (int)(id % pairs);
    const uint64_t trial = trial0 + (uint64_t)b;
    int cnt = 0, cnt_max = 0;
    double c0 = 0.0, c1 = 0.0;
    for (int k = 0; k < T; ++k) {
        if (cnt >= cnt_max) {
            cnt = 0;
            c0 = noise_sample(np, np.rho, np.alpha, trial, (uint32_t)k, (uint32_t)(2 * pr));
            c1 = noise_sample(np, np.rho, np.alpha, trial, (uint32_t)k, (uint32_t)(2 * pr + 1));
            cnt_max = (fabs(c0) > 20.0 || fabs(c1) > 20.0) ? np.hold_cnt : 0;
        } else {
            ++cnt;
        }
        out[((size_t)b * T + k) * m + 2 * pr] = c0;
        out[((size_t)b * T + k) * m + 2 * pr + 1] = c1;
    }
}

cudaError_t launch_noise_fill(const uvs_noise_params& np, unsigned long long trial0, long long B, int T, int m,
                              double* out, cudaStream_t st) {
    const long long n = B * (m / 2);
    noise_fill_kernel<<<(int)((n + 127) / 128), 128, 0, st>>>(np, trial0, B, T, m, out);
    return cudaGetLastError();
}

struct QStart { double q[6]; };

__global__ void start_poses_kernel(QStart qs, unsigned long long pose_seed, int change, unsigned long long trial0,
                                   long long B, double* q_out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double q[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) q[j] = qs.q[j];
    if (change) {                                               // main.py:129-134
        double r0, r1, d;
        uniforms(pose_seed, trial0 + (uint64_t)b, 0, 0, 0, PURPOSE_POSE, r0, d);
        uniforms(pose_seed, trial0 + (uint64_t)b, 0, 1, 0, PURPOSE_POSE, r1, d);
        q[0] = qs.q[0] + 2 * (r0 - 1) * (M_PI / 18);
        q[1] = qs.q[1] + 2 * (r1 - 1) * (M_PI / 9);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) q_out[b * 6 + j] = q[j];
}

cudaError_t launch_start_poses(const double* q_start_host, unsigned long long pose_seed, int change,
                               unsigned long long trial0, long long B, double* q_out, cudaStream_t st) {
    QStart qs;
    for (int j = 0; j < 6; ++j) qs.q[j] = q_start_host[j];
    start_poses_kernel<<<(int)((B + 127) / 128), 128, 0, st>>>(qs, pose_seed, change, trial0, B, q_out);
    return cudaGetLastError();
}

// ---- per-level sums for mean / std (plot_errorbar.m:87-94) ----------------------------------------
__global__ void reduce_stats_kernel(long long B, unsigned long long trial0, int trial_stride, int sweep_epoch, int n_levels,
                                    const double* stats, const int32_t* status, double* sums, double* n_fail) {
    extern __shared__ double sh[];                 // [n_levels][10] : 3 x {n, sx, sxx} + n_fail
    for (int i = threadIdx.x; i < n_levels * 10; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        long long lvl = sweep_epoch > 0 ? (long long)((trial0 + (uint64_t)b * (uint64_t)trial_stride) / (uint64_t)sweep_epoch) : 0;
        if (lvl >= n_levels) lvl = n_levels - 1;
        double* s = sh + lvl * 10;
        if (status[b] == UVS_STATUS_SUCCESS) {
            for (int j = 0; j < 3; ++j) {
                const double x = stats[b * 3 + j];
                atomicAdd(s + 3 * j + 0, 1.0);
                atomicAdd(s + 3 * j + 1, x);
                atomicAdd(s + 3 * j + 2, x * x);
            }
        } else {
            atomicAdd(s + 9, 1.0);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_levels * 10; i += blockDim.x) {
        const int lvl = i / 10, c = i % 10;
        if (sh[i] != 0.0) {
            if (c < 9) atomicAdd(sums + lvl * 9 + c, sh[i]);
            else atomicAdd(n_fail + lvl, sh[i]);
        }
    }
}

cudaError_t launch_reduce_stats(long long B, unsigned long long trial0, int trial_stride, int sweep_epoch, int n_levels,
                                const double* stats, const int32_t* status, double* sums, double* n_fail,
                                cudaStream_t st) {
    int grid = (int)((B + 255) / 256);
    if (grid > 592) grid = 592;
    if (grid < 1) grid = 1;
    reduce_stats_kernel<<<grid, 256, n_levels * 10 * sizeof(double), st>>>(B, trial0, trial_stride > 1 ? trial_stride : 1, sweep_epoch, n_levels, stats,
                                                                           status, sums, n_fail);
    return cudaGetLastError();
}

// ---- FP64 FMA peak probe: 8 independent DFMA chains per thread ------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double* sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.9999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

cudaError_t run_fp64_peak(int iters, int sm_count, double* tflops, double* ms_out) {
    double* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, sizeof(double));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sm_count * 8, block = 256;
    fp64_peak_kernel<<<grid, block>>>(iters / 8 + 1, 1.0, sink);          // warm-up
    cudaEventRecord(e0);
    fp64_peak_kernel<<<grid, block>>>(iters, 1.0, sink);
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)grid * (double)block;
    *tflops = flops / (ms * 1e-3) / 1e12;
    *ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace uvs
