// f1l_pp.cuh -- K1: batched nearest_point + pure-pursuit lookahead on the uploaded track.
//
// Replaces, for B poses at once, utils/utils.py:37-67 (nearest_point), :69-151
// (intersect_point), :153-161 (get_actuation) and pure_pursuit.py:56-122.
//
// Two kernels:
//   pp_scan_kernel    FP32 scan, a grid of ONE-WARP CTAs: a task = (group of 128 poses, part of the
//                     track).  A lane holds four poses as two packed FP32x2 pairs; every lane of
//                     the warp reads the same table entry (a broadcast load, L1-resident: 48 KB
//                     for a 2000-waypoint track).  Why four poses per lane: a broadcast LDS / LDG
//                     still writes 32 x 24 B back to the register file and an SM moves 128 B per
//                     clock, so one pose per lane is bound by that write-back at 24 cycles per
//                     (pose, segment) per warp (measured: l1tex data pipe 87 % busy) -- four
//                     poses per lane cut it to 6 and the packed FFMA2 math (7 FMA-pipe + 1 ALU
//                     instruction per pose and segment, the deviation loop of eval_kernel)
//                     becomes the bound.  Why tasks: 10^5 poses at four per lane are only 782
//                     warps, so the TRACK is split as well; the host picks the number of parts
//                     so that groups x parts fills whole waves of the resident warp slots
//                     (pp_task_parts, occupancy from the CUDA calculator) and the hardware CTA
//                     scheduler balances the one-warp CTAs -- the earlier fixed shape (8-warp
//                     CTAs = 8 parts, CTA barrier, shared-memory combine) ran 2.64 waves at
//                     BASELINE config 2 and left small batches on a fraction of the SMs.  The
//                     parts of a pose meet in a 64-bit atomicMin.  The scan runs on a line form
//                     of each segment in the frame of its 32-segment block (origin kept in
//                     float64), which keeps |coordinates| small where it matters; it only has to
//                     find the right neighbourhood.
//   pp_finish_kernel  one thread per pose.  First the warp re-scans, pose by pose, the winning
//                     32-segment block for the segment index (lanes = segments).  Then float64:
//                     the winner and its +-2 neighbours are re-evaluated operation by operation
//                     like the reference, so index / t / dist / projection agree with the numba
//                     path to rounding; then the lookahead search and get_actuation (or the
//                     Stanley / LQR front-axle errors).
#pragma once
#include "f1l_common.cuh"

#ifndef PP_LANE_POSES
#define PP_LANE_POSES 4        // poses per lane (two packed pairs)
#endif
#define PP_CTA_POSES (32 * PP_LANE_POSES)   // poses of a scan task
#define PP_THREADS 128         // finish kernel

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

// query point of pose `pid`: the pose position, or the front-axle centre (stanley.py:66-68,
// lqr.py:76-78)
__device__ __forceinline__ void pp_query_point(const double* __restrict__ poses, int pose_stride,
                                               int pid, int front_axle, double wb, double& qx,
                                               double& qy, double& qth) {
    qx = poses[(size_t)pose_stride * pid];
    qy = poses[(size_t)pose_stride * pid + 1];
    qth = poses[(size_t)pose_stride * pid + 2];
    if (front_axle) {
        qx = xadd(qx, xmul(wb, cos(qth)));
        qy = xadd(qy, xmul(wb, sin(qth)));
    }
}

// packed pair of track_seg_d2 (segB.y = -uy)
__device__ __forceinline__ f32x2 track_seg_d2_pair(f32x2 px, f32x2 py, float4 A, float2 Bv) {
    const f32x2 ux = pack2(A.x, A.x), uy = pack2(A.y, A.y), nuy = pack2(Bv.y, Bv.y);
    const f32x2 q = ffma2(px, ux, ffma2(py, uy, pack2(A.z, A.z)));
    const f32x2 n = ffma2(py, ux, ffma2(px, nuy, pack2(A.w, A.w)));
    float qa, qb;
    unpack2(q, qa, qb);
    const f32x2 e = pack2(__saturatef(fabsf(qa) + Bv.x), __saturatef(fabsf(qb) + Bv.x));
    return ffma2(e, e, fmul2(n, n));
}

// The scan's copy of the line-form table in CONSTANT memory (tracks of up to F1L_CTAB_SEGS
// segments; one table per device, owned by the handle that uploaded its track last -- others scan
// the global-memory copy).  Every lane of a scan warp reads the same entry, so from constant memory
// the segment constants arrive in UNIFORM registers (LDCU) and enter FFMA2 / FADD.SAT as uniform
// operands (SASS: FFMA2 R14, R6.F32x2.HI_LO, UR13.F32, R12.F32; FADD.SAT R14, |R14|, UR14): they
// no longer cost vector-register-file reads, which is what bounds this loop.
#define F1L_CTAB_SEGS 2560
__constant__ float4 c_segA[F1L_CTAB_SEGS];
__constant__ float2 c_segB[F1L_CTAB_SEGS];

#ifndef PP_TASK_MINB
#define PP_TASK_MINB 24    // resident one-warp scan CTAs per SM asked of the compiler (80 registers)
#endif
#ifndef PP_UNROLL
#define PP_UNROLL 2
#endif

// FP32 scan of the 32-segment blocks [b0, b1) for the PP_LANE_POSES poses of this lane: the
// running minimum of each block (no index bookkeeping in the inner loop) and the best block; the
// exact index is recovered afterwards by re-scanning that one block.
template <bool CT>
__device__ __forceinline__ void pp_scan_blocks(const TrackView& tr, const double (&qx)[PP_LANE_POSES],
                                               const double (&qy)[PP_LANE_POSES], int b0, int b1,
                                               float (&best)[PP_LANE_POSES],
                                               int (&bblk)[PP_LANE_POSES]) {
    const int nseg = tr.n - 1;
#pragma unroll
    for (int p = 0; p < PP_LANE_POSES; ++p) { best[p] = CUDART_INF_F; bblk[p] = 0; }
    // `blk` only ever indexes the table, so that it stays a warp-uniform value to the compiler and
    // the constant-memory entries arrive in uniform registers; everything that needs the block
    // number in a vector register (the origin's address, the per-pose best block) uses an
    // independent copy that the compiler cannot identify with it.
    int blk_v = b0;
    opaque(blk_v);
    for (int blk = b0; blk < b1; ++blk, ++blk_v) {
        const double2 o = tr.blk_origin[blk_v];
        f32x2 px[PP_LANE_POSES / 2], py[PP_LANE_POSES / 2];
#pragma unroll
        for (int p = 0; p < PP_LANE_POSES / 2; ++p) {
            px[p] = pack2((float)(qx[2 * p] - o.x) * TRACK_SCALE, (float)(qx[2 * p + 1] - o.x) * TRACK_SCALE);
            py[p] = pack2((float)(qy[2 * p] - o.y) * TRACK_SCALE, (float)(qy[2 * p + 1] - o.y) * TRACK_SCALE);
        }
        const int k0 = blk << 5;
        const int kn = min(32, nseg - (blk_v << 5));   // the track's last block may be partial
        float m[PP_LANE_POSES];
#pragma unroll
        for (int p = 0; p < PP_LANE_POSES; ++p) m[p] = CUDART_INF_F;
        constexpr int kUnroll = CT ? 2 * PP_UNROLL : PP_UNROLL;   // (constant table: more loads in flight
                                                                  //  per trip cover its cache misses)
        if (CT || kn == 32) {   // (the constant-memory table is padded to whole blocks with far-away entries)
#pragma unroll kUnroll
            for (int j = 0; j < 32; j += 2) {   // two segments per trip, minima by FMNMX3
                const float4 A0 = CT ? c_segA[k0 + j] : __ldg(tr.segA + k0 + j);
                const float4 A1 = CT ? c_segA[k0 + j + 1] : __ldg(tr.segA + k0 + j + 1);
                const float2 B0 = CT ? c_segB[k0 + j] : __ldg(tr.segB + k0 + j);
                const float2 B1 = CT ? c_segB[k0 + j + 1] : __ldg(tr.segB + k0 + j + 1);
#pragma unroll
                for (int p = 0; p < PP_LANE_POSES / 2; ++p) {
                    float a0, a1, c0, c1;
                    unpack2(track_seg_d2_pair(px[p], py[p], A0, B0), a0, a1);
                    unpack2(track_seg_d2_pair(px[p], py[p], A1, B1), c0, c1);
                    m[2 * p] = fmin3(m[2 * p], a0, c0);
                    m[2 * p + 1] = fmin3(m[2 * p + 1], a1, c1);
                }
            }
        } else {
            for (int j = 0; j < kn; ++j) {
                const float4 A0 = CT ? c_segA[k0 + j] : __ldg(tr.segA + k0 + j);
                const float2 B0 = CT ? c_segB[k0 + j] : __ldg(tr.segB + k0 + j);
#pragma unroll
                for (int p = 0; p < PP_LANE_POSES / 2; ++p) {
                    float a0, a1;
                    unpack2(track_seg_d2_pair(px[p], py[p], A0, B0), a0, a1);
                    m[2 * p] = fminf(m[2 * p], a0);
                    m[2 * p + 1] = fminf(m[2 * p + 1], a1);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PP_LANE_POSES; ++p)
            if (m[p] < best[p]) { best[p] = m[p]; bblk[p] = blk_v; }
    }
}

// The 32 lanes evaluate the 32 segments of block `blk` for one query point; the lowest lane
// that attains the block minimum gives the segment index (the first minimum, utils.py:66).
__device__ __forceinline__ int pp_rescan_block(const TrackView& tr, double x, double y, int blk,
                                               int lane) {
    const int nseg = tr.n - 1;
    const double2 o = tr.blk_origin[blk];
    const float prx = (float)(x - o.x) * TRACK_SCALE, pry = (float)(y - o.y) * TRACK_SCALE;
    const int k = (blk << 5) + lane;
    float d2 = CUDART_INF_F;
    if (k < nseg) d2 = track_seg_d2(prx, pry, __ldg(tr.segA + k), __ldg(tr.segB + k));
    const unsigned bits = __float_as_uint(d2);   // d2 >= 0: the bit pattern orders like the value
    const unsigned mn = __reduce_min_sync(F1L_FULL, bits);
    const unsigned hit = __ballot_sync(F1L_FULL, bits == mn);
    return (blk << 5) + __ffs(hit) - 1;
}

// One-warp CTAs, one per (group of 128 poses, part of the track).  The parts of a pose meet in a
// 64-bit atomicMin on (bits(d^2) << 32 | block): the smaller distance wins and, at equal
// distance, the earlier block -- the first minimum in track order (utils.py:66).  `key` is
// preset to all ones; the winning block is re-scanned for the segment index by pp_finish_kernel.
// The track's 32-segment blocks are dealt to the parts as evenly as possible: part p gets
// `per` blocks, the first `rem` parts one more.  Written without an integer division so that a
// task's block range -- and with it every table index of its scan -- is a warp-uniform value to
// the compiler (blockIdx and kernel parameters through uniform-datapath arithmetic only).
#define PP_MAX_PARTS 24
struct PPParts {
    int n, per, rem;
};
template <bool CT>
__global__ void __launch_bounds__(32, PP_TASK_MINB)
pp_scan_kernel(TrackView tr, const double* __restrict__ poses, int pose_stride, int n_poses,
               int front_axle, double wb, PPParts parts, unsigned long long* __restrict__ key) {
    const int lane = threadIdx.x;
    const int grp = blockIdx.x, part = blockIdx.y;   // tasks of one part run together: they walk
                                                     // the same slice of the table
    const int base = grp * PP_CTA_POSES;
    double qx[PP_LANE_POSES], qy[PP_LANE_POSES];
#pragma unroll
    for (int p = 0; p < PP_LANE_POSES; ++p) {
        double qth;
        pp_query_point(poses, pose_stride, min(base + p * 32 + lane, n_poses - 1), front_axle, wb,
                       qx[p], qy[p], qth);
    }
    const int b0 = part * parts.per + min(part, parts.rem);
    const int b1 = b0 + parts.per + (part < parts.rem ? 1 : 0);
    float best[PP_LANE_POSES];
    int bblk[PP_LANE_POSES];
    pp_scan_blocks<CT>(tr, qx, qy, b0, b1, best, bblk);
#pragma unroll
    for (int p = 0; p < PP_LANE_POSES; ++p) {
        const int gid = base + p * 32 + lane;
        if (gid < n_poses)   // (every part holds at least one block: pp_task_parts keeps parts <= blocks)
            atomicMin(key + gid, ((unsigned long long)__float_as_uint(best[p]) << 32) | (unsigned)bblk[p]);
    }
}

// Parts of the track per pose group such that groups x parts one-warp tasks fill whole waves of
// `slots` resident warps: the most efficient count in [6, 24] (fewer, longer tasks on ties); a
// task pays a fixed prologue / epilogue of about 1.5 blocks' worth of work.
// `min_parts`: the constant-memory scan (64 registers, 32 resident warps per SM) measured best
// with 16 or more parts at config 2 (0.0883 ms against 0.0931 with 9), the global-memory one with 6+.
static inline int pp_task_parts(int n_groups, int nblk, int slots, int min_parts = 6) {
    int best_p = 8;
    double best_eff = -1.0;
    if (min_parts > nblk) min_parts = nblk > 0 ? nblk : 1;
    for (int p = min_parts; p <= PP_MAX_PARTS && p <= nblk; ++p) {
        const double w = (double)n_groups * p / slots;
        const double waves = w <= 1.0 ? 1.0 : (double)(long long)(w + 0.999999);
        const double eff = (w / waves) * ((double)nblk / p) / ((double)nblk / p + 1.5);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_p = p; }
    }
    if (best_p > nblk) best_p = nblk > 0 ? nblk : 1;
    return best_p;
}

// `key`: the packed (d^2, block) minima of pp_scan_kernel, one per pose
__global__ void __launch_bounds__(PP_THREADS)
pp_finish_kernel(TrackView tr, const double* __restrict__ poses, int pose_stride, int n_poses, double L, double wb,
                 double max_reacquire, int front_axle, double k_path,
                 const unsigned long long* __restrict__ key, PPOut out) {
    const int nseg = tr.n - 1;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pid = min(gid, n_poses - 1);
    double qx, qy, qth;
    pp_query_point(poses, pose_stride, pid, front_axle, wb, qx, qy, qth);
    // the warp re-scans the winning block of each of its 32 poses, all lanes on one pose at a time
    int bk = 0;
    {
        const int lane = threadIdx.x & 31;
        const int my_blk = (int)(unsigned)(key[pid] & 0xffffffffull);
        const int n_live = min(32, n_poses - (gid - lane));   // warp-uniform
        // four poses per trip: their (load -> distance -> warp minimum -> ballot) chains are
        // independent and overlap (lanes beyond n_live hold the clamped last pose: harmless)
#ifndef PP_RESCAN_UNROLL
#define PP_RESCAN_UNROLL 4
#endif
        for (int i0 = 0; i0 < n_live; i0 += PP_RESCAN_UNROLL) {
            int kk[PP_RESCAN_UNROLL];
#pragma unroll
            for (int u = 0; u < PP_RESCAN_UNROLL; ++u) {
                const int i = i0 + u;
                const int blk = __shfl_sync(F1L_FULL, my_blk, i);
                const double x = __shfl_sync(F1L_FULL, qx, i), y = __shfl_sync(F1L_FULL, qy, i);
                kk[u] = pp_rescan_block(tr, x, y, blk, lane);
            }
#pragma unroll
            for (int u = 0; u < PP_RESCAN_UNROLL; ++u)
                if (lane == i0 + u) bk = kk[u];
        }
    }
    if (gid >= n_poses) return;
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
        ip = intersect_point64(acc, tr.n, qx, qy, L, (double)nr.i + nr.t, true,
                               track_prefilter(tr, qx, qy, L));                  // :71-75
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
