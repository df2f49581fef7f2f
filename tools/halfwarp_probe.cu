// Scratch microbenchmark: does the FP64 pipe of an sm_100a sub-partition (16 lanes per cycle) spend two passes on a warp
// whose upper half is inactive?  One warp per sub-partition (148 blocks x 128 threads), 16 independent DFMA chains per thread,
// lanes >= L exit at once.  If cycles per DFMA halve at L = 16, a small batch (fewer envs than 592 x 32) is better spread as
// 16 envs per warp over twice the warps.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(int iters, int lanes, double* sink, long long* cyc) {
    if ((threadIdx.x & 31) >= lanes) return;
    const double a = 1.0000001, b = 1e-9 * (double)(threadIdx.x + 1);
    double v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 1.0 + 0.1 * j;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fma(v[j], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += v[j];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}
int main() {
    double* sink; long long* cyc;
    const int blocks = 148, iters = 20000;
    for (int warps = 4; warps <= 8; warps += 4) {
        const int threads = 32 * warps;
        cudaMalloc(&sink, blocks * threads * 8); cudaMalloc(&cyc, blocks * warps * 8);
        for (int lanes = 32; lanes >= 4; lanes /= 2) {
            probe<<<blocks, threads>>>(iters, lanes, sink, cyc);
            cudaDeviceSynchronize();
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            probe<<<blocks, threads>>>(iters, lanes, sink, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[148 * 8];
            cudaMemcpy(h, cyc, blocks * warps * 8, cudaMemcpyDeviceToHost);
            double m = 0; for (int i = 0; i < blocks * warps; ++i) m += (double)h[i];
            m /= blocks * warps;
            printf("warps/SM %d  active lanes %2d: %.3f ms, %.2f cycles per DFMA warp-instruction (clock64)\n", warps, lanes, ms,
                   m / (16.0 * iters));
        }
        cudaFree(sink); cudaFree(cyc);
    }
    return 0;
}
