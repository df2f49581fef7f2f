// Scratch microbenchmark: does non-FP64 work issue "for free" beside a saturated FP64 pipe on sm_100a?
// Each thread runs 8 independent DFMA chains; per 8 DFMAs it also executes K independent integer LOP3/IADD ops.
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void probe(int iters, double* sink, unsigned* isink) {
    const double a = 1.0000001, b = 1e-9 * (double)(threadIdx.x + 1);
    double v0 = 1.0, v1 = 1.1, v2 = 1.2, v3 = 1.3, v4 = 1.4, v5 = 1.5, v6 = 1.6, v7 = 1.7;
    unsigned u[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) u[j] = threadIdx.x * 2654435761u + j;
#pragma unroll 2
    for (int it = 0; it < iters; ++it) {
        v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
        v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
#pragma unroll
        for (int j = 0; j < K; ++j) u[j % 16] = (u[j % 16] ^ (u[(j + 5) % 16] >> 3)) + 0x9e3779b9u;   // LOP3/SHF + IADD per op
    }
    unsigned x = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) x ^= u[j];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    isink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int K>
void run(double* sink, unsigned* isink) {
    const int blocks = 148 * 8, threads = 128, iters = 20000;
    probe<K><<<blocks, threads>>>(iters, sink, isink);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<K><<<blocks, threads>>>(iters, sink, isink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = 8.0 * iters * blocks * threads;
    printf("K=%2d int-op groups per 8 DFMA: %.3f ms  %.2f TFLOP/s  (cycles per 8-DFMA group per warp-slot: %.2f)\n", K, ms,
           2 * dfma / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / (iters * (blocks * threads / 32.0) / (148 * 4)));
}
int main() {
    double* sink; unsigned* isink;
    cudaMalloc(&sink, 148 * 8 * 128 * 8); cudaMalloc(&isink, 148 * 8 * 128 * 4);
    run<0>(sink, isink); run<2>(sink, isink); run<4>(sink, isink); run<8>(sink, isink); run<12>(sink, isink); run<16>(sink, isink); run<24>(sink, isink);
    return 0;
}
