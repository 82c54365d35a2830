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

#define PP_CHUNK 512           // segments staged per shared-memory chunk (multiple of 32)
#define PP_THREADS 64
#define PP_SMEM_BYTES (PP_CHUNK * (sizeof(float4) + sizeof(float2)) + (PP_CHUNK / 32) * sizeof(double2))
// Launch shape: 64-thread CTAs with a 12.5 KB table chunk keep ~13 CTAs resident per SM, so the
// 1563 CTAs of 10^5 poses run as ONE balanced wave (21 warps per SM +-5 %); 128-thread CTAs with a
// 48 KB chunk fit 4 per SM and needed 1.3 waves (measured: issue slots busy 48 % of the time).

struct PPOut {
    double* nearest;      // [B,4] proj_x, proj_y, dist, t
    int32_t* nearest_i;   // [B]
    double* lookahead;    // [B,4] p_x, p_y, t2, found
    int32_t* lookahead_i; // [B]
    double* actuation;    // [B,2] steer, speed
    int32_t* status;      // [B]
    double* front;        // [B,6] front-axle mode: theta_e, ef, theta_raceline, kappa_ref,
                          //       goal_velocity, delta (Stanley steering for k_path)
};

// pi_2_pi (utils/utils.py:276-283)
__device__ __forceinline__ double pi_2_pi64(double a) {
    const double pi = 3.14159265358979323846;
    if (a > pi) return a - 2.0 * pi;
    if (a < -pi) return a + 2.0 * pi;
    return a;
}

// FP32 squared distance of a block-relative point to a segment in line form
__device__ __forceinline__ float pp_seg_d2(float prx, float pry, float4 A, float2 Bv) {
    const float q = fmaf(prx, A.x, fmaf(pry, A.y, -A.z));
    const float nn = fmaf(pry, A.x, fmaf(-prx, A.y, -A.w));
    const float t = __saturatef(q * Bv.y);
    const float ex = fmaf(-t, Bv.x, q);
    return fmaf(ex, ex, nn * nn);
}

__global__ void __launch_bounds__(PP_THREADS)
pp_batch_kernel(TrackView tr, const double* __restrict__ poses, int pose_stride, int n_poses, double L, double wb,
                double max_reacquire, int front_axle, double k_path, PPOut out) {
    extern __shared__ __align__(16) unsigned char pp_smem[];
    float4* sA = reinterpret_cast<float4*>(pp_smem);
    float2* sB = reinterpret_cast<float2*>(sA + PP_CHUNK);
    double2* sO = reinterpret_cast<double2*>(sB + PP_CHUNK);

    const int nseg = tr.n - 1;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = gid < n_poses;
    const int pid = active ? gid : n_poses - 1;
    double qx = poses[(size_t)pose_stride * pid], qy = poses[(size_t)pose_stride * pid + 1];
    const double qth = poses[(size_t)pose_stride * pid + 2];
    if (front_axle) {   // query point = front axle centre (stanley.py:66-68, lqr.py:76-78)
        qx = xadd(qx, xmul(wb, cos(qth)));
        qy = xadd(qy, xmul(wb, sin(qth)));
    }

    // The scan keeps only the running minimum of each 32-segment block (one FMNMX per segment
    // instead of compare + two selects) and remembers the best block; the exact index is
    // recovered afterwards by re-scanning that one block (1.6 % extra work).
    float best = CUDART_INF_F;
    int bblk = 0;
    for (int c0 = 0; c0 < nseg; c0 += PP_CHUNK) {
        const int cn = min(PP_CHUNK, nseg - c0);
        const int nblk = (cn + 31) >> 5;
        for (int q = threadIdx.x; q < (nblk << 5); q += blockDim.x) {
            float4 A = make_float4(1.0f, 0.0f, 1e15f, 1e15f);   // padding: far away, finite
            float2 Bv = make_float2(1.0f, 1.0f);
            if (q < cn) {
                A = __ldg(tr.segA + c0 + q);
                Bv = __ldg(tr.segB + c0 + q);
            }
            sA[q] = A;
            sB[q] = Bv;
        }
        for (int b = threadIdx.x; b < nblk; b += blockDim.x) sO[b] = tr.blk_origin[(c0 >> 5) + b];
        __syncthreads();
        for (int blk = 0; blk < nblk; ++blk) {
            const double2 o = sO[blk];
            const float prx = (float)(qx - o.x), pry = (float)(qy - o.y);
            const int j0 = blk << 5;
            float m = CUDART_INF_F;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) m = fminf(m, pp_seg_d2(prx, pry, sA[j0 + j], sB[j0 + j]));
            if (m < best) { best = m; bblk = (c0 >> 5) + blk; }
        }
        __syncthreads();
    }
    if (!active) return;

    // first segment of the best block that attains the block minimum (same FP32 arithmetic)
    int bk = bblk << 5;
    {
        const double2 o = tr.blk_origin[bblk];
        const float prx = (float)(qx - o.x), pry = (float)(qy - o.y);
        float m = CUDART_INF_F;
        const int k1 = min((bblk << 5) + 32, nseg);
        for (int k = bblk << 5; k < k1; ++k) {
            const float d2 = pp_seg_d2(prx, pry, __ldg(tr.segA + k), __ldg(tr.segB + k));
            if (d2 < m) { m = d2; bk = k; }
        }
    }
    // float64 epilogue: exact nearest among the neighbours, then pure_pursuit.py:69-83
    const Nearest64 nr = refine_nearest64(tr.xy, nseg, qx, qy, bk);
    if (front_axle) {
        // stanley.py:69-85 / lqr.py:79-102: cross-track error of the front axle along the
        // vehicle's right-hand normal, heading error against the nearest waypoint's heading
        const double vx = xsub(qx, nr.px), vy = xsub(qy, nr.py);
        const double half_pi = 3.14159265358979323846 / 2.0;
        const double ef = xadd(xmul(vx, cos(qth - half_pi)), xmul(vy, sin(qth - half_pi)));
        const double th_ref = tr.psi[nr.i];
        const double th_e = pi_2_pi64(th_ref - qth);
        const double vel = pose_stride > 3 ? poses[(size_t)pose_stride * pid + 3] : 0.0;
        if (out.front) {
            double* f = out.front + 6 * (size_t)gid;
            f[0] = th_e; f[1] = ef; f[2] = th_ref; f[3] = tr.kappa[nr.i]; f[4] = tr.v[nr.i];
            f[5] = atan2(xmul(k_path, ef), vel) + th_e;   // stanley.py:108-110
        }
        if (out.nearest) {
            double2* o = reinterpret_cast<double2*>(out.nearest + 4 * (size_t)gid);
            o[0] = make_double2(nr.px, nr.py);
            o[1] = make_double2(nr.dist, nr.t);
        }
        if (out.nearest_i) out.nearest_i[gid] = nr.i;
        return;
    }
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
