// Micro-benchmark (B200, sm_100a): is the FP64 tensor-core instruction DMMA.8x8x4 (mma.sync.m8n8k4.f64) worth
// using for the rank-8 covariance update P -= K U^T of the dense RMCKF kernel?
//   (1) DFMA throughput (register-resident, ILP 8), the roofline denominator of bench.py;
//   (2) DMMA throughput with 8 independent accumulators per warp;
//   (3) DMMA dependent-issue latency (one accumulator chain);
//   (4) DFMA and DMMA issued together from the same warps: if the two share one execution pipe the combined
//       flop rate stays at the single-pipe peak, if the tensor path is separate it approaches the sum.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dmma_probe.cu -o tools/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: DFMA only (NF chains); 1: DMMA only (NM accumulators); 2: both interleaved
template <int MODE, int NF, int NM>
__global__ void probe(double* out, long long* cyc, int iters, double a, double b) {
    double x[NF > 0 ? NF : 1], c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) { c0[i] = i; c1[i] = -i; }
    const double fa = 1e-3 * (threadIdx.x & 3), fb = 1e-3 * (threadIdx.x >> 2);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE != 1) {
#pragma unroll
                for (int i = 0; i < NF; ++i) x[i] = fma(x[i], a, b);
            }
            if (MODE != 0) {
#pragma unroll
                for (int i = 0; i < NM; ++i) dmma(c0[i], c1[i], fa, fb);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int NF, int NM>
void run(const char* what, int sms, int warps, double* out, long long* cyc, double ghz) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE, NF, NM><<<sms, warps * 32>>>(out, cyc, 10, 0.999, 1e-3);
    cudaEventRecord(e0);
    probe<MODE, NF, NM><<<sms, warps * 32>>>(out, cyc, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double n_f = (MODE != 1) ? (double)iters * 4 * NF : 0, n_m = (MODE != 0) ? (double)iters * 4 * NM : 0;
    const double flops = (double)sms * warps * (n_f * 64.0 + n_m * 512.0);     // warp DFMA = 32 x 2, warp DMMA = 8 x 8 x 4 x 2
    printf("%-34s warps/SM %2d : %8.3f ms  %6.2f TFLOP/s  | per warp: %6.2f cyc/DFMA %7.2f cyc/DMMA (clock64)\n", what, warps, ms,
           flops / (ms * 1e-3) / 1e12, n_f > 0 ? (double)c / (n_f + n_m) : 0.0, n_m > 0 ? (double)c / (n_m + (MODE == 2 ? n_f : 0)) : 0.0);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s: %d SMs, %.3f GHz\n", p.name, sms, khz * 1e-6);
    double* out; long long* cyc;
    cudaMalloc(&out, (size_t)sms * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int w : {4, 8, 16}) {
        run<0, 8, 0>("DFMA only, 8 chains", sms, w, out, cyc, khz * 1e-6);
        run<1, 0, 8>("DMMA only, 8 accumulators", sms, w, out, cyc, khz * 1e-6);
        run<1, 0, 1>("DMMA only, 1 accumulator (latency)", sms, w, out, cyc, khz * 1e-6);
        run<2, 8, 1>("DFMA x8 + DMMA x1 interleaved", sms, w, out, cyc, khz * 1e-6);
        run<2, 8, 4>("DFMA x8 + DMMA x4 interleaved", sms, w, out, cyc, khz * 1e-6);
    }
    return 0;
}
