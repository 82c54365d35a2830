// mma_bench.cu -- can the raceline-deviation loop take its four FMAs per (sample, segment) from the
// tensor pipe?  q = u.(p - a) - h and n = n^.(p - a) for all (sample, segment) pairs are a rank-2
// affine map: [samples x 8] . [8 x segments] with the operands split into TF32 hi + lo parts
// (x_hi u_hi + x_lo u_hi + x_hi u_lo + x_lo u_lo, products exact in the FP32 accumulator) and the
// per-segment constant entering through the accumulator in full FP32.  The FMA pipe then only does
// e = sat(|q| - h), d^2 = e^2 + n^2 and the running minimum: 3 instead of 6 pipe cycles.
//
//   part A  issue rate of mma.sync.m16n8k8.tf32 (legacy HMMA path) on one SM sub-partition
//   part B  the loop itself: 7 m16 tiles of samples x 128 segments, validated against float64
//           and against the FP32 FFMA formulation of eval_kernel, timed at 4 CTAs x 7 warps / SM
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_bench mma_bench.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <math_constants.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float tf32_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

// D = A (16x8, row) . B (8x8, col) + C, TF32 inputs, FP32 accumulate.  A: a2 = a0 and a3 = a1 here
// (the K columns t and t + 4 of a lane hold the same value), D comes back as two packed pairs:
// (row g: columns 2t, 2t+1), (row g+8: columns 2t, 2t+1).
__device__ __forceinline__ void mma_tf32(f32x2& d01, f32x2& d23, float a0, float a1, float b0, float b1,
                                         float c0, float c1) {
    asm("{\n\t.reg .f32 t0, t1, t2, t3;\n\t"
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {t0, t1, t2, t3}, {%2, %3, %2, %3}, {%4, %5}, {%6, %7, %6, %7};\n\t"
        "mov.b64 %0, {t0, t1};\n\tmov.b64 %1, {t2, t3};\n\t}"
        : "=l"(d01), "=l"(d23)
        : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(b0)),
          "r"(__float_as_uint(b1)), "f"(c0), "f"(c1));
}

#ifndef NWARPS
#define NWARPS 7
#endif
#ifndef MINB
#define MINB 4
#endif
#define UNIT (1.0f / 16384.0f)   // table units (the saturation of e only ever acts at 0)

// ---- part A ------------------------------------------------------------------------------------
template <int ILP>
__global__ void __launch_bounds__(NWARPS * 32, MINB) mma_rate_kernel(float* out, int reps) {
    const int lane = threadIdx.x & 31;
    float a0 = 1.0f + lane * 0.001f, a1 = 0.5f, b0 = 0.25f, b1 = 0.125f;
    f32x2 acc[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i][0] = 0ull; acc[i][1] = 0ull; }
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float c0, c1, c2, c3;
            unpack2(acc[i][0], c0, c1);
            unpack2(acc[i][1], c2, c3);
            asm("{\n\t.reg .f32 t0, t1, t2, t3;\n\t"
                "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {t0, t1, t2, t3}, {%2, %3, %2, %3}, {%4, %5}, {%6, %7, %8, %9};\n\t"
                "mov.b64 %0, {t0, t1};\n\tmov.b64 %1, {t2, t3};\n\t}"
                : "=l"(acc[i][0]), "=l"(acc[i][1])
                : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(b0)),
                  "r"(__float_as_uint(b1)), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float x, y; unpack2(acc[i][0], x, y); s += x + y; unpack2(acc[i][1], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- part B ------------------------------------------------------------------------------------
// geometry in global memory: seg[k] = (ax, ay, bx, by) metres in the ego frame, smp[i] = (x, y)
#define NSEG 128
#define MT 7              // m16 tiles: 112 sample rows
#define NSMP (MT * 16)

struct Tables {
    float4 E[NSEG * 2];   // [seg][parity]: (u_hi, u_lo, -v_hi, -v_lo), (v_hi, v_lo, u_hi, u_lo)
    float4 G[NSEG / 2];   // [seg pair]: (c1a, c1b, c2a, c2b)   c1 = -(a.u + h), c2 = -a.n  (table units)
    float2 H[NSEG / 2];   // [seg pair]: (-h_a, -h_b)
    float4 T[NSEG * 2];   // the FFMA formulation's table (eval_kernel): T0, T1
};

__device__ void build_tables(Tables& tb, const float4* __restrict__ seg) {
    for (int k = threadIdx.x; k < NSEG; k += blockDim.x) {
        const float4 s = seg[k];
        const float dx = s.z - s.x, dy = s.w - s.y;
        const float l2 = fmaf(dx, dx, dy * dy), il = rsqrtf(l2);
        const float ux = dx * il, uy = dy * il, hh = 0.5f * l2 * il;
        const float c1 = -(fmaf(s.x, ux, s.y * uy) + hh) * UNIT;
        const float c2 = -fmaf(s.y, ux, -s.x * uy) * UNIT;
        const float uh = tf32_rna(ux), ul = tf32_rna(ux - uh), vh = tf32_rna(uy), vl = tf32_rna(uy - vh);
        tb.E[2 * k] = make_float4(uh, ul, -vh, -vl);
        tb.E[2 * k + 1] = make_float4(vh, vl, uh, ul);
        float* g = reinterpret_cast<float*>(&tb.G[k >> 1]);
        g[k & 1] = c1;
        g[2 + (k & 1)] = c2;
        reinterpret_cast<float*>(&tb.H[k >> 1])[k & 1] = -hh * UNIT;
        tb.T[2 * k] = make_float4(ux, uy, -uy, il);
        tb.T[2 * k + 1] = make_float4(c1, c2, -hh * UNIT, 0.0f);
    }
}

// MODE 0: tensor-pipe dot products; MODE 1: FP32 FFMA reference formulation, same sample layout
template <int MODE>
__global__ void __launch_bounds__(NWARPS * 32, MINB)
loop_kernel(const float4* __restrict__ seg, const float2* __restrict__ smp, float* __restrict__ dmin,
            float* __restrict__ sink, int reps) {
    __shared__ Tables tb;
    build_tables(tb, seg);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // A fragments: row g / g+8 of every tile; K column t: x_hi, y_hi, x_lo, y_lo
    float a0[MT], a1[MT];
    float px[MT][2], py[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float2 p = smp[m * 16 + g + 8 * r];
            const float x = p.x * UNIT, y = p.y * UNIT;
            px[m][r] = x; py[m][r] = y;
            const float xh = tf32_rna(x), yh = tf32_rna(y);
            const float v = t == 0 ? xh : t == 1 ? yh : t == 2 ? tf32_rna(x - xh) : tf32_rna(y - yh);
            if (r == 0) a0[m] = v; else a1[m] = v;
        }
    }
    float bd[MT][2];
    float acc = 0.f;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int m = 0; m < MT; ++m) { bd[m][0] = CUDART_INF_F; bd[m][1] = CUDART_INF_F; }
        if (MODE == 0) {
#pragma unroll 1
            for (int s0 = 0; s0 < NSEG; s0 += 8) {
                const float4 Bf = tb.E[(s0 + g) * 2 + (t & 1)];
                const float4 Cc = tb.G[(s0 >> 1) + t];
                const float2 nh = tb.H[(s0 >> 1) + t];
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    f32x2 q01, q23, n01, n23;
                    mma_tf32(q01, q23, a0[m], a1[m], Bf.x, Bf.y, Cc.x, Cc.y);
                    mma_tf32(n01, n23, a0[m], a1[m], Bf.z, Bf.w, Cc.z, Cc.w);
                    float qa, qb, qc, qd;
                    unpack2(q01, qa, qb);
                    unpack2(q23, qc, qd);
                    const f32x2 e01 = pack2(__saturatef(fabsf(qa) + nh.x), __saturatef(fabsf(qb) + nh.y));
                    const f32x2 e23 = pack2(__saturatef(fabsf(qc) + nh.x), __saturatef(fabsf(qd) + nh.y));
                    const f32x2 d01 = ffma2(e01, e01, fmul2(n01, n01));
                    const f32x2 d23 = ffma2(e23, e23, fmul2(n23, n23));
                    float da, db, dc, dd;
                    unpack2(d01, da, db);
                    unpack2(d23, dc, dd);
                    bd[m][0] = fmin3(bd[m][0], da, db);
                    bd[m][1] = fmin3(bd[m][1], dc, dd);
                }
            }
        } else {
            // every lane: its 2 rows per tile against segments t, t + 4, ... (same lane-combos)
#pragma unroll 1
            for (int k = t; k < NSEG; k += 8) {
                const float4 A0 = tb.T[2 * k], A1 = tb.T[2 * k + 1];
                const float4 B0 = tb.T[2 * (k + 4)], B1 = tb.T[2 * (k + 4) + 1];
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const f32x2 sx = pack2(px[m][0], px[m][1]), sy = pack2(py[m][0], py[m][1]);
                    f32x2 dA, dB;
                    {
                        const f32x2 q = ffma2(sx, pack2(A0.x, A0.x), ffma2(sy, pack2(A0.y, A0.y), pack2(A1.x, A1.x)));
                        const f32x2 n = ffma2(sy, pack2(A0.x, A0.x), ffma2(sx, pack2(A0.z, A0.z), pack2(A1.y, A1.y)));
                        float qa, qb;
                        unpack2(q, qa, qb);
                        const f32x2 e = pack2(__saturatef(fabsf(qa) + A1.z), __saturatef(fabsf(qb) + A1.z));
                        dA = ffma2(e, e, fmul2(n, n));
                    }
                    {
                        const f32x2 q = ffma2(sx, pack2(B0.x, B0.x), ffma2(sy, pack2(B0.y, B0.y), pack2(B1.x, B1.x)));
                        const f32x2 n = ffma2(sy, pack2(B0.x, B0.x), ffma2(sx, pack2(B0.z, B0.z), pack2(B1.y, B1.y)));
                        float qa, qb;
                        unpack2(q, qa, qb);
                        const f32x2 e = pack2(__saturatef(fabsf(qa) + B1.z), __saturatef(fabsf(qb) + B1.z));
                        dB = ffma2(e, e, fmul2(n, n));
                    }
                    float da, db, ea, eb;
                    unpack2(dA, da, db);
                    unpack2(dB, ea, eb);
                    bd[m][0] = fmin3(bd[m][0], da, ea);
                    bd[m][1] = fmin3(bd[m][1], db, eb);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) acc += bd[m][0] + bd[m][1];
#pragma unroll
        for (int m = 0; m < MT; ++m) { a0[m] += 0.0f * acc; }   // keeps the loop from being hoisted
    }
    // minima over the four lanes of a row group (t = 0..3)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float v = bd[m][r];
            v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 2));
            if (blockIdx.x == 0 && threadIdx.x < 32 && t == 0) dmin[m * 16 + g + 8 * r] = v;
        }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static double ref_dist(const std::vector<float>& seg, double x, double y) {
    double best = 1e300;
    for (int k = 0; k < NSEG; ++k) {
        const double ax = seg[4 * k], ay = seg[4 * k + 1], dx = seg[4 * k + 2] - ax, dy = seg[4 * k + 3] - ay;
        double tt = ((x - ax) * dx + (y - ay) * dy) / (dx * dx + dy * dy);
        tt = tt < 0 ? 0 : tt > 1 ? 1 : tt;
        const double ex = x - (ax + tt * dx), ey = y - (ay + tt * dy);
        best = std::min(best, std::sqrt(ex * ex + ey * ey));
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    const int sms = p.multiProcessorCount;
    const int blocks = sms * MINB * 8;
    printf("%s, %d SMs, %.3f GHz, %d warps x %d CTAs/SM\n", p.name, sms, ghz, NWARPS, MINB);
    float* sink;
    cudaMalloc(&sink, (size_t)blocks * NWARPS * 32 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);

    // part A
    {
        const int reps = 2000;
        auto time_it = [&](auto kern, int ilp, const char* name) {
            float best = 1e30f;
            for (int it = 0; it < 4; ++it) {
                cudaEventRecord(e0);
                kern<<<blocks, NWARPS * 32>>>(sink, reps);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (it && ms < best) best = ms;
            }
            const double n_mma = (double)blocks * NWARPS * reps * ilp;
            const double cyc = best * 1e-3 * ghz * 1e9 * sms * 4;
            printf("mma.sync.m16n8k8.tf32 %-28s %7.3f ms  %6.2f SMSP-cycles per MMA  %7.1f dense TF32 TFLOP/s\n", name, best,
                   cyc / n_mma, n_mma * 2048.0 / (best * 1e-3) / 1e12);
        };
        time_it(mma_rate_kernel<1>, 1, "1 chain / warp");
        time_it(mma_rate_kernel<4>, 4, "4 chains / warp");
        time_it(mma_rate_kernel<8>, 8, "8 chains / warp");
    }

    // part B: a gently curving polyline 128 x 0.2 m starting 6 m behind the ego, 112 samples up to
    // 4 m ahead within +-1.2 m of it
    std::vector<float> seg(4 * NSEG), smp(2 * NSMP);
    {
        double x = -6.0, y = 0.3, th = 0.05;
        for (int k = 0; k < NSEG; ++k) {
            const double nx = x + 0.2 * std::cos(th), ny = y + 0.2 * std::sin(th);
            seg[4 * k] = (float)x; seg[4 * k + 1] = (float)y; seg[4 * k + 2] = (float)nx; seg[4 * k + 3] = (float)ny;
            x = nx; y = ny; th += 0.004 * std::sin(0.05 * k) + 0.003;
        }
        srand(1);
        for (int i = 0; i < NSMP; ++i) {
            const double s = 4.0 * i / (NSMP - 1);
            smp[2 * i] = (float)s;
            smp[2 * i + 1] = (float)(0.3 + 0.1 * s + 1.2 * (rand() / (double)RAND_MAX - 0.5));
        }
    }
    float4* dseg; float2* dsmp; float* dmin;
    cudaMalloc(&dseg, seg.size() * 4); cudaMalloc(&dsmp, smp.size() * 4); cudaMalloc(&dmin, NSMP * 4);
    cudaMemcpy(dseg, seg.data(), seg.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dsmp, smp.data(), smp.size() * 4, cudaMemcpyHostToDevice);
    std::vector<float> got(NSMP);
    for (int mode = 0; mode < 2; ++mode) {
        const int reps = 100;
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            if (mode == 0) loop_kernel<0><<<blocks, NWARPS * 32>>>(dseg, dsmp, dmin, sink, reps);
            else loop_kernel<1><<<blocks, NWARPS * 32>>>(dseg, dsmp, dmin, sink, reps);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it && ms < best) best = ms;
        }
        cudaMemcpy(got.data(), dmin, NSMP * 4, cudaMemcpyDeviceToHost);
        double emax = 0, erel = 0, dsum_ref = 0, dsum = 0;
        for (int i = 0; i < NSMP; ++i) {
            const double r = ref_dist(seg, smp[2 * i], smp[2 * i + 1]);
            const double d = std::sqrt((double)got[i]) * 16384.0;
            emax = std::max(emax, std::fabs(d - r));
            erel = std::max(erel, std::fabs(d - r) / r);
            dsum_ref += r; dsum += d;
        }
        cudaFuncAttributes fa;
        if (mode == 0) cudaFuncGetAttributes(&fa, loop_kernel<0>); else cudaFuncGetAttributes(&fa, loop_kernel<1>);
        const double combos = (double)blocks * NWARPS * reps * (NSEG * NSMP / 32.0);
        const double cyc = best * 1e-3 * ghz * 1e9 * sms * 4;
        printf("%-34s regs %3d spill %3zu B  %7.3f ms  %6.3f SMSP-cycles per (sample, segment)  %6.2f algorithmic TFLOP/s"
               "  | max |d - ref| %.3g m, max rel %.3g, mean-deviation rel err %.3g\n",
               mode == 0 ? "tensor-pipe q/n (TF32 hi+lo), MT=7" : "FP32 FFMA2 q/n (eval_kernel math)", fa.numRegs,
               fa.localSizeBytes, best, cyc / combos, combos * 32 * 17.0 / (best * 1e-3) / 1e12, emax, erel,
               std::fabs(dsum - dsum_ref) / dsum_ref);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
