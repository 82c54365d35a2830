// f1l_peaks.cuh -- FP32 FMA-pipe and MUFU-pipe peak microbenchmarks: the roofline denominators
// for the lattice kernels (MEASURED_PEAKS.json only carries HBM and bf16-GEMM peaks).
#pragma once
#include <cuda_runtime.h>

#define PEAK_CHAINS 16

// PEAK_CHAINS independent register-form FFMA chains per thread
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float b, float c) {
    float a[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) a[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < PEAK_CHAINS; ++k) a[k] = fmaf(a[k], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// PEAK_CHAINS independent MUFU.EX2 chains per thread
__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out, int iters) {
    float a[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) a[k] = (float)(threadIdx.x + k) * 1e-4f - 1.0f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < PEAK_CHAINS; ++k)
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
