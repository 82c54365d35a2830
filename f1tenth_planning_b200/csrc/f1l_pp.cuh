// f1l_pp.cuh -- K1: batched nearest_point + pure-pursuit lookahead on the uploaded track.
//
// Replaces, for B poses at once, utils/utils.py:37-67 (nearest_point), :69-151
// (intersect_point), :153-161 (get_actuation) and pure_pursuit.py:56-122.
//
// Design: one THREAD per pose.  The segment table streams through shared memory in chunks and
// every lane of a warp reads the same segment (a broadcast: one wavefront per LDS), so the
// inner loop is 8 FMA-pipe + 3 ALU instructions per (pose, segment) with no cross-lane traffic.
// The scan runs in FP32 on a line-form of each segment expressed in the frame of its 32-segment
// block (origin kept in float64), which keeps |coordinates| small where it matters; the winner
// and its +-2 neighbours are then re-evaluated in float64 operation by operation like the
// reference, so index / t / dist / projection agree with the numba path to rounding.
#pragma once
#include "f1l_common.cuh"

#define PP_CHUNK 2048          // segments staged per shared-memory chunk (multiple of 32)
#define PP_THREADS 64          // two poses per thread
#define PP_POSES_PER_CTA (2 * PP_THREADS)
#define PP_SMEM_BYTES (PP_CHUNK * 2 * sizeof(float4) + (PP_CHUNK / 32) * sizeof(double2))

struct PPOut {
    double* nearest;      // [B,4] proj_x, proj_y, dist, t
    int32_t* nearest_i;   // [B]
    double* lookahead;    // [B,4] p_x, p_y, t2, found
    int32_t* lookahead_i; // [B]
    double* actuation;    // [B,2] steer, speed
    int32_t* status;      // [B]
};

// FP32 squared distance of a block-relative point to segment k (same operation order as the
// packed scan below, so the bits agree)
__device__ __forceinline__ float pp_seg_d2(float prx, float pry, float4 A, float2 Bv) {
    const float q = fmaf(prx, A.x, fmaf(pry, A.y, -A.z));
    const float nn = fmaf(pry, A.x, fmaf(prx, -A.y, -A.w));
    const float t = __saturatef(q * Bv.y);
    const float e = fmaf(t, -Bv.x, q);
    return fmaf(e, e, nn * nn);
}

// first segment of 32-segment block `blk` whose FP32 distance equals the block minimum
__device__ __forceinline__ int pp_rescan_block(const TrackView& tr, int nseg, double qx, double qy,
                                               int blk) {
    const double2 o = tr.blk_origin[blk];
    const float prx = (float)(qx - o.x), pry = (float)(qy - o.y);
    float best = CUDART_INF_F;
    int bk = blk << 5;
    const int k1 = min((blk << 5) + 32, nseg);
    for (int k = blk << 5; k < k1; ++k) {
        const float d2 = pp_seg_d2(prx, pry, __ldg(tr.segA + k), __ldg(tr.segB + k));
        if (d2 < best) { best = d2; bk = k; }
    }
    return bk;
}

// float64 epilogue of one pose: exact nearest among the FP32 winner's neighbours
// (utils.py:53-66), then pure_pursuit.py:69-83 and get_actuation
__device__ __forceinline__ void pp_epilogue(const TrackView& tr, int nseg, int gid, double qx,
                                            double qy, double qth, int bk, double L, double wb,
                                            double max_reacquire, const PPOut& out) {
    const Nearest64 nr = refine_nearest64(tr.xy, nseg, qx, qy, bk);
    Intersect64 ip;
    ip.px = 0.0; ip.py = 0.0; ip.t = 0.0; ip.i = 0; ip.found = 0;
    int status = 0;
    double lx = 0.0, ly = 0.0, speed = 0.0;
    if (nr.dist < L) {                                            // :70
        XYTrack acc{tr.xy};
        ip = intersect_point64(acc, tr.n, qx, qy, L, (double)nr.i + nr.t, true);  // :71-75
        if (ip.found) {                                           // :78
            const int r = pymod(ip.i, tr.n);
            lx = tr.xy[r].x; ly = tr.xy[r].y; speed = tr.v[nr.i];
            status = 1;
        }
    } else if (nr.dist < max_reacquire) {                         // :80-81
        lx = tr.xy[nr.i].x; ly = tr.xy[nr.i].y; speed = tr.v[nr.i];
        status = 2;
    }
    double steer = 0.0;
    if (status) steer = actuation_steer64(qth, lx, ly, qx, qy, L, wb);  // :116-120
    else speed = 0.0;                                                   // :112-114

    if (out.nearest) {
        double2* o = reinterpret_cast<double2*>(out.nearest + 4 * (size_t)gid);
        o[0] = make_double2(nr.px, nr.py);
        o[1] = make_double2(nr.dist, nr.t);
    }
    if (out.nearest_i) out.nearest_i[gid] = nr.i;
    if (out.lookahead) {
        double2* o = reinterpret_cast<double2*>(out.lookahead + 4 * (size_t)gid);
        o[0] = make_double2(ip.px, ip.py);
        o[1] = make_double2(ip.t, ip.found ? 1.0 : 0.0);
    }
    if (out.lookahead_i) out.lookahead_i[gid] = ip.found ? ip.i : 0;
    if (out.actuation)
        *reinterpret_cast<double2*>(out.actuation + 2 * (size_t)gid) = make_double2(steer, speed);
    if (out.status) out.status[gid] = status;
}


__global__ void __launch_bounds__(PP_THREADS)
pp_batch_kernel(TrackView tr, const double* __restrict__ poses, int pose_stride, int n_poses, double L, double wb,
                double max_reacquire, PPOut out) {
    extern __shared__ __align__(16) unsigned char pp_smem[];
    // per segment two float4: T0 = (ux, uy, -uy, 1/len), T1 = (-a.u, -a.n, -len, 0)
    float4* sT = reinterpret_cast<float4*>(pp_smem);
    double2* sO = reinterpret_cast<double2*>(sT + 2 * PP_CHUNK);

    const int nseg = tr.n - 1;
    const int g0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int pa = min(g0, n_poses - 1), pb = min(g0 + 1, n_poses - 1);
    const double qxa = poses[(size_t)pose_stride * pa], qya = poses[(size_t)pose_stride * pa + 1];
    const double qxb = poses[(size_t)pose_stride * pb], qyb = poses[(size_t)pose_stride * pb + 1];

    // two poses per thread in packed FP32x2 registers: the scan keeps only the running minimum
    // (one FMNMX per pose and segment) and the 32-segment block it came from; the exact index
    // is recovered afterwards by re-scanning that one block.
    float besta = CUDART_INF_F, bestb = CUDART_INF_F;
    int blka = 0, blkb = 0;
    for (int c0 = 0; c0 < nseg; c0 += PP_CHUNK) {
        const int cn = min(PP_CHUNK, nseg - c0);
        const int nblk = (cn + 31) >> 5;
        for (int q = threadIdx.x; q < (nblk << 5); q += blockDim.x) {
            float4 T0 = make_float4(1.0f, 0.0f, -0.0f, 1.0f);      // padding: far away, finite
            float4 T1 = make_float4(-1e15f, -1e15f, -1.0f, 0.0f);
            if (q < cn) {
                const float4 A = __ldg(tr.segA + c0 + q);
                const float2 Bv = __ldg(tr.segB + c0 + q);
                T0 = make_float4(A.x, A.y, -A.y, Bv.y);
                T1 = make_float4(-A.z, -A.w, -Bv.x, 0.0f);
            }
            sT[2 * q] = T0;
            sT[2 * q + 1] = T1;
        }
        for (int b = threadIdx.x; b < nblk; b += blockDim.x) sO[b] = tr.blk_origin[(c0 >> 5) + b];
        __syncthreads();
        for (int blk = 0; blk < nblk; ++blk) {
            const double2 o = sO[blk];
            const f32x2 prx = pack2((float)(qxa - o.x), (float)(qxb - o.x));
            const f32x2 pry = pack2((float)(qya - o.y), (float)(qyb - o.y));
            const float4* T = sT + 2 * (blk << 5);
            float ma = CUDART_INF_F, mb = CUDART_INF_F;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                const float4 T0 = T[2 * j], T1 = T[2 * j + 1];
                const f32x2 q2 = ffma2(prx, pack2(T0.x, T0.x), ffma2(pry, pack2(T0.y, T0.y), pack2(T1.x, T1.x)));
                const f32x2 n2 = ffma2(pry, pack2(T0.x, T0.x), ffma2(prx, pack2(T0.z, T0.z), pack2(T1.y, T1.y)));
                float qa, qb;
                unpack2(q2, qa, qb);
                const f32x2 t2 = pack2(__saturatef(qa * T0.w), __saturatef(qb * T0.w));
                const f32x2 e2 = ffma2(t2, pack2(T1.z, T1.z), q2);
                const f32x2 d2 = ffma2(e2, e2, fmul2(n2, n2));
                float da, db;
                unpack2(d2, da, db);
                ma = fminf(ma, da);
                mb = fminf(mb, db);
            }
            if (ma < besta) { besta = ma; blka = (c0 >> 5) + blk; }
            if (mb < bestb) { bestb = mb; blkb = (c0 >> 5) + blk; }
        }
        __syncthreads();
    }

    for (int half = 0; half < 2; ++half) {
        const int gid = g0 + half;
        if (gid >= n_poses) return;
        const double qx = half ? qxb : qxa, qy = half ? qyb : qya;
        const double qth = poses[(size_t)pose_stride * gid + 2];
        const int bk = pp_rescan_block(tr, nseg, qx, qy, half ? blkb : blka);
        pp_epilogue(tr, nseg, gid, qx, qy, qth, bk, L, wb, max_reacquire, out);
    }
}

// intersect_point for independent queries (API parity for the free function)
__global__ void intersect_batch_kernel(TrackView tr, const double* __restrict__ pts,
                                       const double* __restrict__ t0, int n, double radius,
                                       int wrap, double* __restrict__ out,
                                       int32_t* __restrict__ out_i) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    XYTrack acc{tr.xy};
    const Intersect64 ip = intersect_point64(acc, tr.n, pts[2 * (size_t)gid],
                                             pts[2 * (size_t)gid + 1], radius, t0[gid], wrap != 0);
    out[4 * (size_t)gid] = ip.px;
    out[4 * (size_t)gid + 1] = ip.py;
    out[4 * (size_t)gid + 2] = ip.t;
    out[4 * (size_t)gid + 3] = ip.found ? 1.0 : 0.0;
    out_i[gid] = ip.found ? ip.i : 0;
}

// get_actuation for independent queries; in [n,7], out [n,2] = (speed, steer)
__global__ void actuation_batch_kernel(const double* __restrict__ in, int n, double wb,
                                       double* __restrict__ out) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    const double* r = in + 7 * (size_t)gid;
    out[2 * (size_t)gid] = r[3];
    out[2 * (size_t)gid + 1] = actuation_steer64(r[0], r[1], r[2], r[4], r[5], r[6], wb);
}
