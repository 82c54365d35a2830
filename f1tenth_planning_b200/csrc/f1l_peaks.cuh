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

// packed FP32x2 FMA (Blackwell FFMA2): PEAK_CHAINS independent float2 chains per thread
__global__ void __launch_bounds__(256) ffma2_peak_kernel(float* out, int iters, float b, float c) {
    float2 a[PEAK_CHAINS];
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) a[k] = make_float2((float)(threadIdx.x + k) * 1e-3f, (float)k * 1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < PEAK_CHAINS; ++k) a[k] = __ffma2_rn(a[k], b2, c2);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) s += a[k].x + a[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// issue-slot probe: per iteration 8 FFMA + 8 FMNMX (ALU pipe) per thread, all independent chains:
// tells whether an FMA-pipe instruction and an ALU-pipe instruction share the issue slot
__global__ void __launch_bounds__(256) mixed_peak_kernel(float* out, int iters, float b, float c) {
    float a[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = (float)(threadIdx.x + k) * 1e-3f; m[k] = 1e9f - k; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a[k] = fmaf(a[k], b, c);
            m[k] = fminf(m[k], a[(k + 3) & 7]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + m[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
