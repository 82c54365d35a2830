// f1l_lattice_ws.cuh -- eval_ws_kernel: the fused generate / cost / collision kernel of the batch
// regime as a PRODUCER / CONSUMER pipeline inside eight-warp CTAs (one CTA per scenario).
//
// Why: the raceline-deviation loop is 63 % of eval_kernel's instructions and its schedule is held
// back by the 72-register cap that 28 resident warps per SM impose -- in isolation the loop runs
// at 9.65 sub-partition cycles per (sample, segment) at 72 registers and at 8.18 at 92
// (tools/microbench/devloop_bench.cu), while the other stages (goals, Newton, samples, curvature,
// collision) are latency-bound, need few registers and want many warps.  One register budget for
// both loses either way (profiles/r2_microbench.md).  Here the two kinds of work run in different
// warps of the same CTA with different budgets (setmaxnreg):
//
//   warps 0-3  producers (56 registers): take 4-candidate items of the scenario, solve / sample /
//              check them exactly like eval_kernel, and hand every valid candidate -- its samples
//              in the pair layout of the deviation pass plus a small record -- to their consumer
//              through a two-slot ring in shared memory (mbarrier full / empty);
//   warps 4-7  consumers (104 registers): load a slot's samples into registers, release the slot,
//              run the segment loop with a whole trip's temporaries in flight, reduce, finish the
//              cost and enter it into the scenario's argmin.
//
// The arithmetic of every stage is eval_kernel's, statement by statement: costs, flags and the
// argmin are bit-identical (tests/test_gpu_lattice.py::test_ws_kernel_*).  Used for batches of
// whole-scenario CTAs with the cubic generator, M <= 104, prune_window = 0 and the work counters
// off; everything else runs eval_kernel.
#pragma once
#include "f1l_lattice.cuh"

#ifndef WS_PRODUCERS
#define WS_PRODUCERS 8
#endif
#define WS_CONSUMERS 4
#define WS_WARPS (WS_PRODUCERS + WS_CONSUMERS)
#ifndef WS_SLOTS
#define WS_SLOTS 16   // depth of the CTA's ring (a power of two)
#endif
#ifndef WS_REGS_PRODUCER
#define WS_REGS_PRODUCER 56
#endif
#ifndef WS_REGS_CONSUMER
#define WS_REGS_CONSUMER 104
#endif
// what the CTA is launched with: the pool the two setmaxnreg redistribute
#define WS_REGS_LAUNCH ((WS_PRODUCERS * WS_REGS_PRODUCER + WS_CONSUMERS * WS_REGS_CONSUMER) / WS_WARPS)
#define WS_MINB (WS_PRODUCERS == 8 ? 2 : 3)

#define WS_STR2(x) #x
#define WS_STR(x) WS_STR2(x)

// Ring slots are handed back and forth through two monotonic use counters per slot -- `filled`
// (uses the producers have completed) and `released` (uses the consumers have taken out) -- not
// through phase-parity barriers: with several producers and consumers on one ring a slow producer
// can be lapped twice, and a parity cannot tell use k from use k + 2.  A ticket t names slot
// t % WS_SLOTS and use t / WS_SLOTS; its holder is the only one who advances that counter to its
// value, so the waits are plain equality tests (acquire loads, release stores, a short sleep
// between polls).
__device__ __forceinline__ void ws_seq_store(uint32_t addr, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void ws_seq_wait(uint32_t addr, uint32_t want) {
    uint32_t v;
    for (;;) {
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        if (v == want) break;
        __nanosleep(40);
    }
}

// Shared-memory layout (bytes; ws_smem_bytes on the host mirrors it).  Per producer: probe list |
// item solutions; per producer / consumer pair: WS_SLOTS x (record | slab x | slab y); then the
// barriers and the CTA-wide part of eval_kernel (opponents | grid constants | previous path |
// window table).
template <int S, int SG>
struct WsSmem {
    static constexpr int PCAP = S * SG;
    static constexpr int SROWS = (S + 1) / 2;
    static constexpr int SLAB = SROWS * SG * 2;                      // floats per coordinate
    static constexpr uint32_t P_PLIST = 0;                           // [PCAP][2] float4
    static constexpr uint32_t P_ITEM = P_PLIST + PCAP * 32;          // [4][2] float4
    static constexpr uint32_t P_BYTES = P_ITEM + 128;
    static constexpr uint32_t SLOT_REC = 0;                          // 2 float4
    static constexpr uint32_t SLOT_X = 32;
    static constexpr uint32_t SLOT_Y = SLOT_X + SLAB * 4;
    static constexpr uint32_t SLOT_BYTES = SLOT_Y + SLAB * 4;
    static constexpr uint32_t RING = WS_PRODUCERS * P_BYTES;         // [slot]
    static constexpr uint32_t BAR_FULL = RING + WS_SLOTS * SLOT_BYTES;   // [slot] u64
    static constexpr uint32_t BAR_EMPTY = BAR_FULL + WS_SLOTS * 8;
    static constexpr uint32_t C_OPP = BAR_EMPTY + WS_SLOTS * 8;
    static constexpr uint32_t C_GRID = C_OPP + F1L_MAX_OPP * 16;
    static constexpr uint32_t C_PREV = C_GRID + 48;
    static constexpr uint32_t C_TAB = C_PREV + F1L_MAX_M * 4;
    static_assert(P_BYTES % 16 == 0 && SLOT_BYTES % 16 == 0 && C_OPP % 16 == 0, "16-byte alignment");
};

template <int IPL, int S, int SG>
__global__ void __launch_bounds__(WS_WARPS * 32, WS_MINB) eval_ws_kernel(EvalArgs a) {
    constexpr int GG = 32 / SG;
    using L = WsSmem<S, SG>;
    extern __shared__ __align__(16) unsigned char ev_smem[];
    // the CTA's ring is multi-producer / multi-consumer: write and read tickets, the next item of
    // the scenario, producers that have run out of items
    __shared__ unsigned int s_wticket, s_rticket, s_done;
    __shared__ int s_next;
    const int M = a.ep.M;
    const int ntab = a.nseg_pad + EVAL_SEG_PAD;
    constexpr int PCAP = L::PCAP;
    constexpr int SLAB = L::SLAB;
    constexpr int NT = WS_WARPS * 32;

    const int s = blockIdx.x;                 // one CTA per scenario
    const int tid = threadIdx.x, wid = tid >> 5;
    int lane = tid & 31;
    uint32_t cbase = smem_u32(ev_smem);
    opaque(lane);
    opaque(cbase);
    const QueryCtx* __restrict__ q = a.ctx + s;
    const int cb = a.c_begin, ce = a.c_end;

    // ---- prologue (all eight warps): raceline window -> vehicle frame -> line form (eval_kernel's) ----
    {
        const double px = q->px, py = q->py;
        const float cth = q->cth, sth = q->sth;
        const int seg0 = q->seg0, nseg = q->nseg, ns = a.tr.n - 1;
        for (int k = tid; k < ntab; k += NT) {
            float4 T0 = make_float4(1.0f, 0.0f, -0.0f, 1.0f);      // padding: far away, finite
            float4 T1 = make_float4(-1e9f, -1e9f, -1.0f, 0.0f);    // (table units) d^2 = 1e18
            if (k < nseg) {
                int sg = seg0 + k;
                if (sg >= ns) sg -= ns;
                const double2 p0 = a.tr.xy[sg], p1 = a.tr.xy[sg + 1];
                const float ax = (float)(p0.x - px), ay = (float)(p0.y - py);
                const float dxm = (float)(p1.x - p0.x), dym = (float)(p1.y - p0.y);
                const float avx = fmaf(cth, ax, sth * ay), avy = fmaf(cth, ay, -sth * ax);
                const float dx = fmaf(cth, dxm, sth * dym), dy = fmaf(cth, dym, -sth * dxm);
                const float l2 = fmaf(dx, dx, dy * dy);
                const float il = rsqrtf(l2);
                const float ux = dx * il, uy = dy * il;
                T0 = make_float4(ux, uy, -uy, il);
                const float hh = 0.5f * (l2 * il);
                T1 = make_float4(-(fmaf(avx, ux, avy * uy) + hh) * EVAL_DEV_SCALE,
                                 -fmaf(avy, ux, -avx * uy) * EVAL_DEV_SCALE, -hh * EVAL_DEV_SCALE, 0.0f);
            }
            sts128(cbase + L::C_TAB + k * 32, T0);
            sts128(cbase + L::C_TAB + k * 32 + 16, T1);
        }
        if (a.prev_theta) {
#pragma unroll 1
            for (int i = tid; i < M; i += NT) sts32(cbase + L::C_PREV + i * 4, a.prev_theta[i]);
        }
        if (tid < F1L_MAX_OPP)
            sts128(cbase + L::C_OPP + tid * 16,
                   tid < q->n_opp ? q->opp[tid] : make_float4(1e9f, 1e9f, 1.0f, 0.0f));
        if (tid == 32) {
            sts128(cbase + L::C_GRID, make_float4(q->gA00, q->gA01, q->gA10, q->gA11));
            sts128(cbase + L::C_GRID + 16, make_float4(q->gfx, q->gfy, __int_as_float(q->gix), __int_as_float(q->giy)));
            sts128(cbase + L::C_GRID + 32, make_float4(__int_as_float(q->n_opp), __int_as_float(q->has_grid), 0.0f, 0.0f));
        }
        if (tid < WS_SLOTS) {   // use counters of the ring slots
            sts32(cbase + L::BAR_FULL + tid * 8, 0.0f);
            sts32(cbase + L::BAR_EMPTY + tid * 8, 0.0f);
        }
        if (tid == 0) { s_wticket = 0u; s_rticket = 0u; s_done = 0u; s_next = cb + WS_PRODUCERS * 4; }
        // unused slots of the last sample rows of every ring slot: finite dummies, written once
        for (int i = tid; i < WS_SLOTS * (S * SG - M); i += NT) {
            const int slot = i / (S * SG - M), e = M + (i - slot * (S * SG - M));
            const int r = e / SG, g = e - r * SG;
            const uint32_t at = (uint32_t)(((r >> 1) * SG + g) * 2 + (r & 1)) * 4;
            const uint32_t sb = cbase + L::RING + slot * L::SLOT_BYTES;
            sts32(sb + L::SLOT_X + at, 0.0f);
            sts32(sb + L::SLOT_Y + at, 0.0f);
        }
    }
    __syncthreads();

    if (wid < WS_PRODUCERS) {
        // =============================== producer ===============================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " WS_STR(WS_REGS_PRODUCER) ";");
        uint32_t wbase = cbase + wid * L::P_BYTES;
        const uint32_t ring = cbase + L::RING;
        const uint32_t bar_full = cbase + L::BAR_FULL;
        const uint32_t bar_empty = cbase + L::BAR_EMPTY;
        int nown = min(max(M - lane * IPL, 0), IPL);
        opaque(wbase);
        opaque(nown);
        int slot = 0;
        uint32_t use = 0;
        bool have_slot = false;    // a write ticket is held: its slot is empty and not yet handed over

        // producers pull the scenario's 4-candidate items from the CTA's counter
        const int p_hi = ce;
        for (int c0 = cb + wid * 4; c0 < ce;) {
            {   // goals, LUT seeds and Newton solves of the item's four candidates, one per 8-lane group
                const int cg = c0 + (lane >> 3);
                const bool active = cg < p_hi;
                float ggx, ggy, ggth, gp3, gv_ref;
                bool ghave;
                candidate_goal(a.centres, a.widths, a.goals, a.nL, a.nW, a.inv_nW, a.C, s, active ? cg : c0,
                               a.ep.use_goal_kappa != 0, ggx, ggy, ggth, gp3, ghave, gv_ref);
                SpiralF gsp;
                const int g_pass = generate_cubic_g8(gsp, a.lut, a.ep, ggx, ggy, ggth, gp3, lane, active);
                if ((lane & 7) == 0) {
                    const uint32_t ia = wbase + L::P_ITEM + 32 * (lane >> 3);
                    sts128(ia, make_float4(ggx, ggy, ggth, gp3));
                    sts128(ia + 16, make_float4(gsp.p1, gsp.p2, gsp.sf, __int_as_float((ghave ? 256 : 0) | g_pass)));
                }
                __syncwarp();
            }
            const int c_last = min(c0 + 4, p_hi);
            for (int c = c0; c < c_last; ++c) {
                float gx, gy, gth, p3;
                bool have_centre;
                SpiralF sp;
                int n_pass;
                {
                    const uint32_t ia = wbase + L::P_ITEM + 32 * (c - c0);
                    const float4 G = lds128(ia), Q = lds128(ia + 16);
                    gx = G.x; gy = G.y; gth = G.z; p3 = G.w;
                    const int hp = __float_as_int(Q.w);
                    have_centre = (hp & 256) != 0;
                    n_pass = hp & 255;
                    sp.p0 = 0.0f;
                    sp.p3 = p3;
                    sp.p1 = Q.x; sp.p2 = Q.y; sp.sf = Q.z;
                    spiral_set(sp);
                }
                float x[IPL], y[IPL], th[IPL], kp[IPL], cs[IPL], sn[IPL];
                spiral_sample<IPL>(sp, M, lane, x, y, th, kp, cs, sn);

                const size_t cand = (size_t)s * a.C + c;
                if (a.states) {
#pragma unroll
                    for (int j = 0; j < IPL; ++j)
                        if (j < nown) a.states[cand * M + lane * IPL + j] = make_float4(x[j], y[j], th[j], kp[j]);
                }
                if (a.headings) {
#pragma unroll
                    for (int j = 0; j < IPL; ++j)
                        if (j < nown) a.headings[cand * M + lane * IPL + j] = make_float2(cs[j], sn[j]);
                }

                // the slot the samples go to must be empty before the slab is written
                if (!have_slot) {   // next write ticket of the ring; its slot must have been released
                    unsigned int tk = 0;
                    if (lane == 0) tk = atomicAdd(&s_wticket, 1u);
                    tk = __shfl_sync(F1L_FULL, tk, 0);
                    slot = (int)(tk % WS_SLOTS);
                    use = tk / WS_SLOTS;
                    ws_seq_wait(bar_empty + slot * 8, use);   // every earlier use has been taken out
                    have_slot = true;
                }
                const uint32_t sb = ring + slot * L::SLOT_BYTES;

                float maxk = 0.0f, sumk = 0.0f, ex = 0.0f, ey = 0.0f, eth = 0.0f;
                {
                    const int i0 = lane * IPL;
                    const int r0 = i0 / SG, g0 = i0 - r0 * SG;
                    const uint32_t wa = sb + L::SLOT_X + (uint32_t)(((r0 >> 1) * SG + g0) * 2 + (r0 & 1)) * 4;
                    const int last_lane = (M - 1) / IPL, last_j = (M - 1) - last_lane * IPL;
#pragma unroll
                    for (int j = 0; j < IPL; ++j) {
                        if (j < nown) {
                            const float ak = fabsf(kp[j]);
                            maxk = fmaxf(maxk, ak);
                            sumk += ak;
                            uint32_t at = wa + 8 * j;
                            if constexpr (SG % IPL != 0) {
                                const int i = i0 + j, r = i / SG, g = i - r * SG;
                                at = sb + L::SLOT_X + (uint32_t)(((r >> 1) * SG + g) * 2 + (r & 1)) * 4;
                            }
                            sts32(at, x[j] * EVAL_DEV_SCALE);
                            sts32(at + SLAB * 4, y[j] * EVAL_DEV_SCALE);
                        }
                        if (j == last_j) { ex = x[j]; ey = y[j]; eth = th[j]; }
                    }
                    ex = __shfl_sync(F1L_FULL, ex, last_lane);
                    ey = __shfl_sync(F1L_FULL, ey, last_lane);
                    eth = __shfl_sync(F1L_FULL, eth, last_lane);
                }
                maxk = warp_max(maxk);
                sumk = warp_sum(sumk);
                const float gn = sqrtf(fmaf(gx, gx, fmaf(gy, gy, gth * gth)));
                const float tol = a.ep.tol * fmaxf(gn, 1.0f);
                bool valid = have_centre && isfinite(sp.p1) && isfinite(sp.p2) && isfinite(sp.sf) &&
                             sp.sf > 0.0f && fabsf(ex - gx) < tol && fabsf(ey - gy) < tol &&
                             fabsf(eth - gth) < tol;
                if (valid && a.ep.kappa_max > 0.0f && !(maxk <= a.ep.kappa_max)) valid = false;

                unsigned flags = valid ? F1L_FLAG_VALID : 0u;
                if (!have_centre) flags |= F1L_FLAG_NO_CENTRE;
                flags |= (unsigned)min(n_pass, 15) << F1L_FLAG_PASS_SHIFT;
                if (lane == 0) {
                    if (a.goals_out) {
                        float* g = a.goals_out + cand * 3;
                        g[0] = gx; g[1] = gy; g[2] = gth;
                    }
                    if (a.params) a.params[cand] = make_float4(sp.p1, sp.p2, sp.sf, sp.p3);
                }

                if (valid) {  // warp-uniform
                    const float t_len = __fdividef(1.0f, sp.sf);
                    const float t_maxk = maxk;
                    const float t_meank = sumk * a.ep.inv_M;
                    float t_sim = 0.0f;
                    float sim = 0.0f;
                    bool hit_opp = false, hit_map = false;
                    const uint8_t* occ = a.grid.occ;
                    const int gw = a.grid.w, gh = a.grid.h;
                    const float4 GA = lds128(cbase + L::C_GRID), GB = lds128(cbase + L::C_GRID + 16);
                    const float4 GC = lds128(cbase + L::C_GRID + 32);
                    const int n_opp = __float_as_int(GC.x);
                    const bool has_grid = __float_as_int(GC.y) != 0;
                    const float hl = a.ep.half_l, hw = a.ep.half_w;
                    const float A00 = GA.x, A01 = GA.y, A10 = GA.z, A11 = GA.w;
                    const float gfx = GB.x, gfy = GB.y;
                    const int gix = __float_as_int(GB.z), giy = __float_as_int(GB.w);
                    int clr[IPL];
#pragma unroll
                    for (int j = 0; j < IPL; ++j) {
                        clr[j] = 0;
                        if (has_grid && a.grid.near_map && j < nown) {
                            const float ccx = ffm(A00, x[j], ffm(A01, y[j], gfx));
                            const float ccy = ffm(A10, x[j], ffm(A11, y[j], gfy));
                            const int ccol = gix + __float2int_rd(ccx), crow = giy + __float2int_rd(ccy);
                            if ((unsigned)ccol < (unsigned)gw && (unsigned)crow < (unsigned)gh)
                                clr[j] = __ldg(a.grid.near_map + (size_t)crow * gw + ccol);
                        }
                    }
                    unsigned opp_mask;
                    {
                        const float4 o = lds128(cbase + L::C_OPP + (lane & (F1L_MAX_OPP - 1)) * 16);
                        const float mx = o.x - 0.5f * ex, my = o.y - 0.5f * ey;
                        const float reach = 0.5f * sp.sf + a.ep.reach_pad;
                        opp_mask = __ballot_sync(F1L_FULL, lane < n_opp && fmaf(mx, mx, my * my) <= reach * reach);
                    }
                    if (a.prev_theta) {
                        const int lim = M - a.ep.n_shift - a.ep.n_cull;
#pragma unroll
                        for (int j = 0; j < IPL; ++j) {
                            const int i = lane * IPL + j;
                            if (i < lim) {
                                const float d = th[j] - lds32(cbase + L::C_PREV + (i + a.ep.n_shift) * 4);
                                sim = fmaf(d, d, sim);
                            }
                        }
                    }
                    float lane_R2 = -1.0f;
                    if (opp_mask && nown > 0) {
                        const float R = fmaf(1.01f * (float)(IPL - 1), __fdividef(sp.sf, (float)(M - 1)), a.ep.reach_pad);
                        lane_R2 = R * R;
                    }
                    for (unsigned m = opp_mask; m; m &= m - 1) {
                        const float4 o = lds128(cbase + L::C_OPP + (__ffs(m) - 1) * 16);
                        const float ux = o.x - x[0], uy = o.y - y[0];
                        if (fmaf(ux, ux, uy * uy) <= lane_R2) {
#pragma unroll
                            for (int j = 0; j < IPL; ++j) {
                                if (j < nown) {
                                    const float tx = fs(o.x, x[j]), ty = fs(o.y, y[j]);
                                    const float d2 = fa(fm(tx, tx), fm(ty, ty));
                                    if (d2 <= a.ep.rc2 && sat_collide(tx, ty, cs[j], sn[j], o.z, o.w, hl, hw))
                                        hit_opp = true;
                                }
                            }
                        }
                    }
                    if (has_grid) {
                        unsigned nm[IPL];
                        int total = 0;
#pragma unroll
                        for (int j = 0; j < IPL; ++j) {
                            nm[j] = __ballot_sync(F1L_FULL, j < nown && clr[j] < a.grid.near_free);
                            total += __popc(nm[j]);
                        }
                        if (total) {
                            const uint32_t pl = wbase + L::P_PLIST;
                            int rank0 = 0;
                            const bool discs = a.ep.collision_mode == 1;
                            const float al = discs ? a.grid.disc_off : hl;
#pragma unroll
                            for (int j = 0; j < IPL; ++j) {
                                if ((nm[j] >> lane) & 1u) {
                                    const int r = rank0 + __popc(nm[j] & ((1u << lane) - 1u));
                                    const float ccx = ffm(A00, x[j], ffm(A01, y[j], gfx));
                                    const float ccy = ffm(A10, x[j], ffm(A11, y[j], gfy));
                                    const float lx = fm(cs[j], al), ly = fm(sn[j], al);
                                    const float wx = fm(-sn[j], hw), wy = fm(cs[j], hw);
                                    sts128(pl + 32 * r, make_float4(ccx, ccy, fa(fm(A00, lx), fm(A01, ly)),
                                                                    fa(fm(A10, lx), fm(A11, ly))));
                                    sts128(pl + 32 * r + 16, make_float4(fa(fm(A00, wx), fm(A01, wy)),
                                                                         fa(fm(A10, wx), fm(A11, wy)), 0.0f, 0.0f));
                                }
                                rank0 += __popc(nm[j]);
                            }
                            __syncwarp();
                            if (discs) {
                                const int nwork = total * 3;
                                for (int w = lane; w < nwork; w += 32) {
                                    const int fp = w / 3;
                                    const float sa = (float)(w - 3 * fp) - 1.0f;
                                    const float4 P0 = lds128(pl + 32 * fp);
                                    const int col = gix + __float2int_rd(fa(P0.x, fm(sa, P0.z)));
                                    const int row = giy + __float2int_rd(fa(P0.y, fm(sa, P0.w)));
                                    if ((unsigned)col >= (unsigned)gw || (unsigned)row >= (unsigned)gh) hit_map = true;
                                    else if ((int)__ldg(a.grid.edt2 + (size_t)row * gw + col) < a.grid.disc_t2) hit_map = true;
                                }
                            } else {
                                const int nwork = total * 9;
                                for (int w = lane; w < nwork; w += 32) {
                                    const int fp = w / 9, p = w - 9 * fp;
                                    const float4 P0 = lds128(pl + 32 * fp), P1 = lds128(pl + 32 * fp + 16);
                                    const float sa = (float)(int)((0x1520au >> (2 * p)) & 3u) - 1.0f;
                                    const float sb2 = (float)(int)((0x12522u >> (2 * p)) & 3u) - 1.0f;
                                    hit_map |= grid_hit(occ, gw, gh, gix, giy, fa(fa(P0.x, fm(sa, P0.z)), fm(sb2, P1.x)),
                                                        fa(fa(P0.y, fm(sa, P0.w)), fm(sb2, P1.y)));
                                }
                            }
                            __syncwarp();
                        }
                    }
                    if (a.prev_theta) t_sim = warp_sum(sim);
                    hit_opp = __any_sync(F1L_FULL, hit_opp);
                    hit_map = __any_sync(F1L_FULL, hit_map);
                    if (hit_opp) flags |= F1L_FLAG_COLLIDE_OPP;
                    if (hit_map) flags |= F1L_FLAG_COLLIDE_MAP;
                    const float partial = a.ep.w[0] * t_len + a.ep.w[1] * t_maxk + a.ep.w[2] * t_meank + a.ep.w[3] * t_sim;
                    if (lane == 0) {
                        if (a.terms) {
                            float* t = a.terms + cand * F1L_N_TERMS;
                            t[0] = t_len; t[1] = t_maxk; t[2] = t_meank; t[3] = t_sim;
                        }
                        if (a.flags) a.flags[cand] = (uint8_t)flags;
                        // hand-over: record (candidate, flags, partial cost), samples already in the slab
                        sts128(sb + L::SLOT_REC, make_float4(__int_as_float(c), __int_as_float((int)flags), partial, 0.0f));
                    }
                    __syncwarp();
                    if (lane == 0) ws_seq_store(bar_full + slot * 8, use + 1);
                    have_slot = false;
                } else {
                    // failed validation: the candidate is finished here; its ticket goes through the
                    // ring as "skip" (-2) right away, so that nobody queues behind an unfilled slot
                    if (lane == 0) {
                        if (a.costs) a.costs[cand] = CUDART_INF_F;
                        if (a.flags) a.flags[cand] = (uint8_t)flags;
                        if (a.terms) {
                            float* t = a.terms + cand * F1L_N_TERMS;
                            t[0] = 0.0f; t[1] = 0.0f; t[2] = 0.0f; t[3] = 0.0f; t[4] = 0.0f;
                        }
                        atomicMin(a.best + s, ((unsigned long long)float_orderable(CUDART_INF_F) << 32) | (unsigned)c);
                        sts128(sb + L::SLOT_REC, make_float4(__int_as_float(-2), 0.0f, 0.0f, 0.0f));
                        ws_seq_store(bar_full + slot * 8, use + 1);
                    }
                    __syncwarp();
                    have_slot = false;
                }
            }
            __syncwarp();
            int c_next = 0;
            if (lane == 0) c_next = atomicAdd(&s_next, 4);
            c0 = __shfl_sync(F1L_FULL, c_next, 0);
        }
        // the last producer to run out of items closes the ring with one "end" record (-1) per consumer
        unsigned int fin = 0;
        if (lane == 0) { __threadfence_block(); fin = atomicAdd(&s_done, 1u); }
        fin = __shfl_sync(F1L_FULL, fin, 0);
        if (fin == WS_PRODUCERS - 1) {
            for (int e = 0; e < WS_CONSUMERS; ++e) {
                unsigned int tk = 0;
                if (lane == 0) tk = atomicAdd(&s_wticket, 1u);
                tk = __shfl_sync(F1L_FULL, tk, 0);
                const int sl = (int)(tk % WS_SLOTS);
                ws_seq_wait(bar_empty + sl * 8, tk / WS_SLOTS);
                if (lane == 0) {
                    sts128(ring + sl * L::SLOT_BYTES + L::SLOT_REC, make_float4(__int_as_float(-1), 0.0f, 0.0f, 0.0f));
                    ws_seq_store(bar_full + sl * 8, tk / WS_SLOTS + 1);
                }
                __syncwarp();
            }
        }
    } else {
        // =============================== consumer ===============================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " WS_STR(WS_REGS_CONSUMER) ";");
        const uint32_t ring = cbase + L::RING;
        const uint32_t bar_full = cbase + L::BAR_FULL;
        const uint32_t bar_empty = cbase + L::BAR_EMPTY;
        const int sgi = lane / GG, ggi = lane - sgi * GG;
        const int nrows = min(max((M - sgi + SG - 1) / SG, 0), S);
        constexpr int SP = S / 2;
        constexpr bool ODD = (S & 1) != 0;
        const bool split_odd = ODD && (M - (S - 1) * SG) <= SG / 2;   // uniform
        const int nq = a.nseg_pad;
        for (;;) {
            unsigned int tk = 0;
            if (lane == 0) tk = atomicAdd(&s_rticket, 1u);
            tk = __shfl_sync(F1L_FULL, tk, 0);
            const int slot = (int)(tk % WS_SLOTS);
            ws_seq_wait(bar_full + slot * 8, tk / WS_SLOTS + 1);
            const uint32_t sb = ring + slot * L::SLOT_BYTES;
            const float4 rec = lds128(sb + L::SLOT_REC);
            const int c = __float_as_int(rec.x);
            if (c == -1) break;          // end of the ring for this consumer
            if (c < 0) {                 // a ticket without a candidate: release the slot and go on
                __syncwarp();
                if (lane == 0) ws_seq_store(bar_empty + slot * 8, tk / WS_SLOTS + 1);
                continue;
            }
            const unsigned flags = (unsigned)__float_as_int(rec.y);
            const float partial = rec.z;

            f32x2 sx2[SP], sy2[SP];
            float bdx[SP], bdy[SP];
            float sxl = 0.0f, syl = 0.0f, bdl = CUDART_INF_F;
            const uint32_t ra = sb + L::SLOT_X + sgi * 8;
#pragma unroll
            for (int j = 0; j < SP; ++j) {
                sx2[j] = lds64(ra + j * SG * 8);
                sy2[j] = lds64(ra + SLAB * 4 + j * SG * 8);
                bdx[j] = CUDART_INF_F;
                bdy[j] = CUDART_INF_F;
            }
            if (ODD) {
                const uint32_t ro = split_odd ? sb + L::SLOT_X + (sgi & (SG / 2 - 1)) * 8 : ra;
                sxl = lds32(ro + SP * SG * 8);
                syl = lds32(ro + SLAB * 4 + SP * SG * 8);
            }
            // the samples are in registers: the producer may refill the slot while the loop runs
            __syncwarp();
            if (lane == 0) ws_seq_store(bar_empty + slot * 8, tk / WS_SLOTS + 1);

            uint32_t ta = cbase + L::C_TAB + ggi * 32;
            float4 T0 = lds128(ta), T1 = lds128(ta + 16);
            if (split_odd) {
                const uint32_t off = (sgi >= SG / 2) ? GG * 32 : 0;
                for (int n = nq / (2 * GG); n > 0; --n) {
                    const float4 B0 = lds128(ta + GG * 32), B1 = lds128(ta + GG * 32 + 16);
                    seg_min2<SP>(sx2, sy2, T0, T1, B0, B1, bdx, bdy);
                    const uint32_t to = ta + off;
                    bdl = fminf(bdl, seg_dist2(sxl, syl, lds128(to), lds128(to + 16)));
                    ta += 2 * GG * 32;
                    T0 = lds128(ta);
                    T1 = lds128(ta + 16);
                }
                bdl = fminf(bdl, __shfl_xor_sync(F1L_FULL, bdl, (SG / 2) * GG));
            } else {
                for (int n = nq / (2 * GG); n > 0; --n) {
                    const float4 B0 = lds128(ta + GG * 32), B1 = lds128(ta + GG * 32 + 16);
                    seg_min2<SP>(sx2, sy2, T0, T1, B0, B1, bdx, bdy);
                    if (ODD) bdl = fmin3(bdl, seg_dist2(sxl, syl, T0, T1), seg_dist2(sxl, syl, B0, B1));
                    ta += 2 * GG * 32;
                    T0 = lds128(ta);
                    T1 = lds128(ta + 16);
                }
            }
            float rows[S];
#pragma unroll
            for (int j = 0; j < SP; ++j) { rows[2 * j] = bdx[j]; rows[2 * j + 1] = bdy[j]; }
            if (ODD) rows[S - 1] = bdl;
            const float dsum = halve_min_sqrt_sum<S, GG>(rows, ggi, nrows);
            const float t_dev = warp_sum(dsum) * ((1.0f / EVAL_DEV_SCALE) * a.ep.inv_M);

            float cost = CUDART_INF_F;
            if (!(flags & (F1L_FLAG_COLLIDE_OPP | F1L_FLAG_COLLIDE_MAP))) {
                cost = partial + a.ep.w[4] * t_dev;
                if (!isfinite(cost)) cost = CUDART_INF_F;
            }
            if (lane == 0) {
                const size_t cand = (size_t)s * a.C + c;
                if (a.costs) a.costs[cand] = cost;
                if (a.terms) a.terms[cand * F1L_N_TERMS + 4] = t_dev;
                atomicMin(a.best + s, ((unsigned long long)float_orderable(cost) << 32) | (unsigned)c);
            }
        }
    }
}
