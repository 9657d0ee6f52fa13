// Micro-benchmark: FP64 DFMA dependent-issue latency and per-SM throughput vs warps/ILP (B200).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dfma_lat.cu -o tools/dfma_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps, double* out, long long* cyc) {
    const int iters = 2000;
    k<ILP><<<148, warps * 32>>>(out, cyc, iters, 0.999, 1e-3);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / (iters * 8.0 * ILP);            // cycles per warp-DFMA as seen by one warp
    double sm_rate = warps * 32.0 / per;                        // thread-DFMA / cycle / SM
    printf("ILP %d warps/SM %2d : %.2f cyc per DFMA per warp, %.1f DFMA lanes/cycle/SM\n", ILP, warps, per, sm_rate);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
    return 0;
}
