// rf_bench.cu -- how fast do FFMA / FFMA2 / FMNMX issue as a function of how many of their source
// operands come from distinct registers (register-file read bandwidth), per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o rf_bench rf_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
#define N 8
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// variant v:
// 0 FFMA  a = a*B + C          B, C kernel parameters (constant bank): 1 register read
// 1 FFMA  a = a*x[k] + C       2 register reads
// 2 FFMA  a = x[k]*y[k] + a    3 register reads
// 3 FFMA2 a2 = a2*b2 + c2      b2, c2 loop-invariant register pairs (reuse cache candidates)
// 4 FFMA2 a2 = x2[k]*y2[k] + a2   three distinct register pairs (6 words)
// 5 FFMA2 a2 = x2[k]*y[k].F32 + a2   pair, scalar broadcast, pair (5 words)
// 6 FFMA2 a2 = x2[k]*y[k].F32 + z[k].F32   (4 words), result feeds nothing else (dependent only through x2 rotation)
// 7 FMNMX m = min(m, x[k])     2 register reads
// 8 FMNMX3 m = max3(-x[k], y[k], m)
// 9 mix: per k one 3-register FFMA + one FMNMX (2 reads)
// 10 mix: per k one FFMA2 (5 words) + one FMNMX3 (3 words)
template <int V>
__global__ void __launch_bounds__(256) k(float* out, int iters, float B, float C) {
    float a[N], x[N], y[N], z[N], m[N];
    f32x2 a2[N], x2[N], y2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float* in = out + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0 + 64 * i;   // opaque inputs
        a[i] = in[threadIdx.x & 7]; x[i] = in[8 + (threadIdx.x & 7)]; y[i] = in[16 + (threadIdx.x & 7)];
        z[i] = in[24 + (threadIdx.x & 7)]; m[i] = in[32 + (threadIdx.x & 7)];
        a2[i] = pack2(a[i], a[i] + 1.f); x2[i] = pack2(x[i], x[i] * 0.999f); y2[i] = pack2(y[i], y[i] * 2.f);
    }
    const f32x2 b2 = pack2(B, B * 1.0001f), c2 = pack2(C, C * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (V == 0) a[i] = fmaf(a[i], B, C);
            if (V == 1) a[i] = fmaf(a[i], x[i], C);
            if (V == 2) a[i] = fmaf(x[i], y[i], a[i]);
            if (V == 3) a2[i] = ffma2(a2[i], b2, c2);
            if (V == 4) a2[i] = ffma2(x2[i], y2[i], a2[i]);
            if (V == 5) a2[i] = ffma2(x2[i], pack2(y[i], y[i]), a2[i]);
            if (V == 6) a2[i] = ffma2(a2[i], pack2(y[i], y[i]), pack2(z[i], z[i]));
            if (V == 7) m[i] = fminf(m[i], x[i]);
            if (V == 8) m[i] = fmax3(-x[i], y[i], m[i]);
            if (V == 9) { a[i] = fmaf(x[i], y[i], a[i]); m[i] = fminf(m[i], z[i]); }
            if (V == 10) { a2[i] = ffma2(x2[i], pack2(y[i], y[i]), a2[i]); m[i] = fmax3(-x[i], z[i], m[i]); }
        }
        if (V == 7 || V == 8 || V == 9 || V == 10) {   // keep the min/max chains from collapsing
#pragma unroll
            for (int i = 0; i < N; ++i) if (it == iters + 5) { x[i] += 1.f; z[i] -= 1.f; }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) { float lo, hi; unpack2(a2[i], lo, hi); s += a[i] + m[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(const char* name, float* out, int sms, double ghz, int inst_per_k) {
    const int iters = 4096, blocks = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<V><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    const double warp_inst = (double)blocks * 8 * iters * N * inst_per_k;
    const double cyc = best * 1e-3 * ghz * 1e9 * sms * 4;
    printf("%-58s %7.3f ms  %.3f warp-instructions / clk / SMSP\n", name, best, warp_inst / cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    float* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4); cudaMemset(out, 0, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    const int sms = p.multiProcessorCount;
    run<0>("FFMA  a*C0+C1 (1 register read)", out, sms, ghz, 1);
    run<1>("FFMA  a*x+C (2 register reads)", out, sms, ghz, 1);
    run<2>("FFMA  x*y+a (3 register reads)", out, sms, ghz, 1);
    run<3>("FFMA2 a2*b2+c2 (loop-invariant pairs)", out, sms, ghz, 1);
    run<4>("FFMA2 x2*y2+a2 (3 distinct pairs, 6 words)", out, sms, ghz, 1);
    run<5>("FFMA2 x2*y.F32+a2 (5 words)", out, sms, ghz, 1);
    run<6>("FFMA2 a2*y.F32+z.F32 (4 words)", out, sms, ghz, 1);
    run<7>("FMNMX m=min(m,x)", out, sms, ghz, 1);
    run<8>("FMNMX3 m=max3(-x,y,m)", out, sms, ghz, 1);
    run<9>("FFMA(3 reads)+FMNMX(2 reads)", out, sms, ghz, 2);
    run<10>("FFMA2(5 words)+FMNMX3(3 words)", out, sms, ghz, 2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
