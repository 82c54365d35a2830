// devloop_bench.cu -- the raceline-deviation segment loop of eval_kernel in isolation, in several
// formulations: every resident warp runs the loop only, so the result is the loop's own ceiling
// (SM sub-partition cycles per sample x segment) without the other stages of the kernel mixed in.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o devloop_bench devloop_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

#ifndef NWARPS
#define NWARPS 7
#endif
#ifndef MINB
#define MINB 4
#endif

// MATH 0: e = max(-q, q - len, 0)  (FADD2 + 2 FMNMX3)      T1 = (-a.u, -a.n, -len)
// MATH 1: e = sat(|q_c| - h), q_c = q - h, h = len / 2, coordinates in units of 64 m so that the
//         saturation never clips (2 scalar FADD.SAT with |.| modifier, no ALU-pipe instruction)
//                                                           T1 = (-(a.u + h), -a.n, -h)
template <int MATH>
__device__ __forceinline__ f32x2 dist2_pair(f32x2 sx, f32x2 sy, const float4& T0, const float4& T1) {
    const f32x2 ux = pack2(T0.x, T0.x), uy = pack2(T0.y, T0.y);
    const f32x2 nuy = pack2(T0.z, T0.z), nc = pack2(T1.x, T1.x);
    const f32x2 ne = pack2(T1.y, T1.y), nlen = pack2(T1.z, T1.z);
    const f32x2 q2 = ffma2(sx, ux, ffma2(sy, uy, nc));
    const f32x2 n2 = ffma2(sy, ux, ffma2(sx, nuy, ne));
    float qa, qb;
    unpack2(q2, qa, qb);
    f32x2 e2;
    if (MATH == 0) {
        float ra, rb;
        unpack2(fadd2(q2, nlen), ra, rb);
        e2 = pack2(fmax3(-qa, ra, 0.0f), fmax3(-qb, rb, 0.0f));
    } else {
        e2 = pack2(__saturatef(fabsf(qa) + T1.z), __saturatef(fabsf(qb) + T1.z));
    }
    return ffma2(e2, e2, fmul2(n2, n2));
}

// ORDER 0: pair after pair; ORDER 1: operation-major over groups of 3 pairs (consecutive
// instructions share their segment-constant operands: register reuse cache)
// MIN3 0: one segment per trip, FMNMX; 1: two segments per trip, one FMNMX3 per sample
template <int SP, int MATH, int ORDER>
__device__ __forceinline__ void seg_step(const f32x2 (&sx2)[SP], const f32x2 (&sy2)[SP], const float4& T0,
                                         const float4& T1, f32x2 (&d)[SP]) {
    if (ORDER == 0) {
#pragma unroll
        for (int j = 0; j < SP; ++j) d[j] = dist2_pair<MATH>(sx2[j], sy2[j], T0, T1);
    } else {
        const f32x2 ux = pack2(T0.x, T0.x), uy = pack2(T0.y, T0.y);
        const f32x2 nuy = pack2(T0.z, T0.z), nc = pack2(T1.x, T1.x);
        const f32x2 ne = pack2(T1.y, T1.y), nlen = pack2(T1.z, T1.z);
#pragma unroll
        for (int j0 = 0; j0 < SP; j0 += 3) {
            f32x2 q[3], n[3], e[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j0 + j < SP) q[j] = ffma2(sy2[j0 + j], uy, nc);
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j0 + j < SP) n[j] = ffma2(sx2[j0 + j], nuy, ne);
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j0 + j < SP) q[j] = ffma2(sx2[j0 + j], ux, q[j]);
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j0 + j < SP) n[j] = ffma2(sy2[j0 + j], ux, n[j]);
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j0 + j < SP) {
                float qa, qb;
                unpack2(q[j], qa, qb);
                if (MATH == 0) {
                    float ra, rb;
                    unpack2(fadd2(q[j], nlen), ra, rb);
                    e[j] = pack2(fmax3(-qa, ra, 0.0f), fmax3(-qb, rb, 0.0f));
                } else {
                    e[j] = pack2(__saturatef(fabsf(qa) + T1.z), __saturatef(fabsf(qb) + T1.z));
                }
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j0 + j < SP) n[j] = fmul2(n[j], n[j]);
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j0 + j < SP) d[j0 + j] = ffma2(e[j], e[j], n[j]);
        }
    }
}

template <int S, int SG, int MATH, int ORDER, int MIN3>
__global__ void __launch_bounds__(NWARPS * 32, MINB) loop_kernel(float* out, int reps, int nq) {
    constexpr int GG = 32 / SG, SP = S / 2;
    static_assert((S & 1) == 0, "even S only");
    __shared__ float4 sT[2 * (256 + 32)];
    for (int k = threadIdx.x; k < 256 + 32; k += blockDim.x) {
        const float ang = 0.01f * k;
        sT[2 * k] = make_float4(cosf(ang), sinf(ang), -sinf(ang), 5.0f);
        sT[2 * k + 1] = make_float4(-0.002f * k, 0.0001f * k, -0.002f, 0.0f);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int sgi = lane / GG, ggi = lane - sgi * GG;
    f32x2 sx2[SP], sy2[SP];
    float bdx[SP], bdy[SP];
#pragma unroll
    for (int j = 0; j < SP; ++j) {
        sx2[j] = pack2(out[j + lane], out[2 * j + lane]);
        sy2[j] = pack2(out[3 * j + lane], out[j + 2 * lane]);
        bdx[j] = CUDART_INF_F;
        bdy[j] = CUDART_INF_F;
    }
    for (int r = 0; r < reps; ++r) {
        float4 T0 = sT[2 * ggi], T1 = sT[2 * ggi + 1];
        if (MIN3 == 0) {
#pragma unroll 2
            for (int k = ggi; k < nq; k += GG) {
                const float4 N0 = sT[2 * (k + GG)];
                const float4 N1 = sT[2 * (k + GG) + 1];
                f32x2 d[SP];
                seg_step<SP, MATH, ORDER>(sx2, sy2, T0, T1, d);
#pragma unroll
                for (int j = 0; j < SP; ++j) {
                    float da, db;
                    unpack2(d[j], da, db);
                    bdx[j] = fminf(bdx[j], da);
                    bdy[j] = fminf(bdy[j], db);
                }
                T0 = N0;
                T1 = N1;
            }
        } else {
            for (int k = ggi; k < nq; k += 2 * GG) {
                const float4 B0 = sT[2 * (k + GG)], B1 = sT[2 * (k + GG) + 1];
                f32x2 d[SP], e[SP];
                seg_step<SP, MATH, ORDER>(sx2, sy2, T0, T1, d);
                seg_step<SP, MATH, ORDER>(sx2, sy2, B0, B1, e);
#pragma unroll
                for (int j = 0; j < SP; ++j) {
                    float da, db, ea, eb;
                    unpack2(d[j], da, db);
                    unpack2(e[j], ea, eb);
                    bdx[j] = fmin3(bdx[j], da, ea);
                    bdy[j] = fmin3(bdy[j], db, eb);
                }
                T0 = sT[2 * (k + 2 * GG)];
                T1 = sT[2 * (k + 2 * GG) + 1];
            }
        }
#pragma unroll
        for (int j = 0; j < SP; ++j) sx2[j] = fadd2(sx2[j], pack2(1e-3f, 1e-3f));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < SP; ++j) s += bdx[j] + bdy[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The same loop with every lane of the warp on the SAME segment (lanes = sample groups only) and
// the segment constants in CONSTANT memory: the index is warp-uniform, the entries arrive in uniform
// registers (LDCU) and enter FFMA2 / FADD.SAT as uniform operands -- no vector-register-file reads
// for them.  This is what an "item-wide" eval loop (the 4 x 100 samples of an item on one warp
// against the track's own table) or K1's scan look like.  Two segments per trip, FMNMX3 minima.
__constant__ float4 cT[2 * (256 + 32)];
template <int S>
__global__ void __launch_bounds__(NWARPS * 32, MINB) loop_kernel_ct(float* out, int reps, int nq) {
    constexpr int SP = S / 2;
    const int lane = threadIdx.x & 31;
    f32x2 sx2[SP], sy2[SP];
    float bdx[SP], bdy[SP];
#pragma unroll
    for (int j = 0; j < SP; ++j) {
        sx2[j] = pack2(out[j + lane], out[2 * j + lane]);
        sy2[j] = pack2(out[3 * j + lane], out[j + 2 * lane]);
        bdx[j] = CUDART_INF_F;
        bdy[j] = CUDART_INF_F;
    }
    for (int r = 0; r < reps; ++r) {
        for (int k = 0; k < nq; k += 2) {
            const float4 T0 = cT[2 * k], T1 = cT[2 * k + 1], B0 = cT[2 * k + 2], B1 = cT[2 * k + 3];
            f32x2 d[SP], e[SP];
            seg_step<SP, 1, 0>(sx2, sy2, T0, T1, d);
            seg_step<SP, 1, 0>(sx2, sy2, B0, B1, e);
#pragma unroll
            for (int j = 0; j < SP; ++j) {
                float da, db, ea, eb;
                unpack2(d[j], da, db);
                unpack2(e[j], ea, eb);
                bdx[j] = fmin3(bdx[j], da, ea);
                bdy[j] = fmin3(bdy[j], db, eb);
            }
        }
#pragma unroll
        for (int j = 0; j < SP; ++j) sx2[j] = fadd2(sx2[j], pack2(1e-3f, 1e-3f));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < SP; ++j) s += bdx[j] + bdy[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int S>
void run_ct(const char* name, float* out, int sms, double ghz) {
    const int reps = 25, nq = 128;
    const int blocks = sms * MINB * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        loop_kernel_ct<S><<<blocks, NWARPS * 32>>>(out, reps, nq);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
    }
    const double warp_sample_seg = (double)blocks * NWARPS * reps * nq * S;
    const double smsp_cycles = best * 1e-3 * ghz * 1e9 * sms * 4;
    const double flops = warp_sample_seg * 32 * 17.0;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, loop_kernel_ct<S>);
    printf("%-44s regs %3d spill %3zu B  %7.3f ms  %6.3f SMSP-cycles per (sample, segment)  %6.2f algorithmic TFLOP/s\n",
           name, fa.numRegs, fa.localSizeBytes, best, smsp_cycles / warp_sample_seg, flops / (best * 1e-3) / 1e12);
}

template <int S, int SG, int MATH, int ORDER, int MIN3>
void run(const char* name, float* out, int sms, double ghz) {
    const int reps = 100, nq = 128;
    const int blocks = sms * MINB * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        loop_kernel<S, SG, MATH, ORDER, MIN3><<<blocks, NWARPS * 32>>>(out, reps, nq);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
    }
    constexpr int GG = 32 / SG;
    const double warp_sample_seg = (double)blocks * NWARPS * reps * (nq / GG) * S;
    const double smsp_cycles = best * 1e-3 * ghz * 1e9 * sms * 4;
    const double flops = warp_sample_seg * 32 * 17.0;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, loop_kernel<S, SG, MATH, ORDER, MIN3>);
    printf("%-44s regs %3d spill %3zu B  %7.3f ms  %6.3f SMSP-cycles per (sample, segment)  %6.2f algorithmic TFLOP/s\n",
           name, fa.numRegs, fa.localSizeBytes, best, smsp_cycles / warp_sample_seg, flops / (best * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    float* out;
    const size_t nb = (size_t)p.multiProcessorCount * MINB * 8 * NWARPS * 32 * sizeof(float);
    cudaMalloc(&out, nb);
    cudaMemset(out, 0, nb);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.3f GHz, %d warps x %d CTAs/SM\n", p.name, sms, ghz, NWARPS, MINB);
    run<12, 8, 0, 0, 0>("S=12 max3 math, pair order", out, sms, ghz);
    run<12, 8, 1, 0, 0>("S=12 FADD.SAT math, pair order", out, sms, ghz);
    run<12, 8, 0, 1, 0>("S=12 max3 math, op-major x3", out, sms, ghz);
    run<12, 8, 1, 1, 0>("S=12 FADD.SAT math, op-major x3", out, sms, ghz);
    run<12, 8, 1, 0, 1>("S=12 FADD.SAT math, pair order, min3", out, sms, ghz);
    run<12, 8, 1, 1, 1>("S=12 FADD.SAT math, op-major x3, min3", out, sms, ghz);
    run<8, 16, 1, 0, 0>("S=8 SG=16 FADD.SAT math, pair order", out, sms, ghz);
    run<8, 16, 1, 1, 1>("S=8 SG=16 FADD.SAT, op-major x3, min3", out, sms, ghz);
    run<6, 16, 1, 1, 1>("S=6 SG=16 FADD.SAT, op-major x3, min3", out, sms, ghz);
    {
        float4 hT[2 * (256 + 32)];
        for (int k = 0; k < 256 + 32; ++k) {
            const float ang = 0.01f * k;
            hT[2 * k] = make_float4(cosf(ang), sinf(ang), -sinf(ang), 5.0f);
            hT[2 * k + 1] = make_float4(-0.002f * k, 0.0001f * k, -0.002f, 0.0f);
        }
        cudaMemcpyToSymbol(cT, hT, sizeof(hT));
    }
    run_ct<12>("S=12 uniform operands (constant table), min3", out, sms, ghz);
    run_ct<14>("S=14 uniform operands (constant table), min3", out, sms, ghz);
    run_ct<8>("S=8 uniform operands (constant table), min3", out, sms, ghz);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
