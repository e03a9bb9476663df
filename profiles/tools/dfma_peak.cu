// FP64 FMA peak of the device (denominator for the FP64-pipe-bound forward stencil).
// 8 independent DFMA chains per thread, 1024 threads per SM-resident CTA, 2 CTAs per SM.
// build + run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma profiles/tools/dfma_peak.cu && /tmp/dfma
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 2, threads = 1024, iters = 4096;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
    printf("%s: %d SMs, DFMA peak %.2f TFLOP/s fp64 (%.3f ms, best of 5)\n", p.name, p.multiProcessorCount,
           flop / best / 1e9, best);
    return cudaGetLastError() != cudaSuccess;
}
