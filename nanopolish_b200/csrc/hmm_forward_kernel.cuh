// hmm_forward_kernel.cuh — K1: the R9 profile-HMM forward score on sm_100a (kernel template).
//
// Replaces, for a whole batch of (sequence, HMMInputData, flags) jobs at once:
//   profile_hmm_score_r9            ref: src/hmm/nanopolish_profile_hmm_r9.cpp:35-65
//   profile_hmm_fill_generic_r9     ref: src/hmm/nanopolish_profile_hmm_r9.inl:265-433
//   ProfileHMMForwardOutputR9       ref: src/hmm/nanopolish_profile_hmm_r9.inl:79-127
//   log_probability_match_r9        ref: src/hmm/nanopolish_emissions.h:57-68
//   get_scaled_gaussian_from_pore_model_state   ref: src/nanopolish_squiggle_read.h:217-226
//   p7_FLogsum                      ref: src/common/logsum.h:55-66
//
// Design (DESIGN.md section 3 has the long form):
//   * a group of W lanes (W = 4, 8, 16 or 32; 32/W jobs per warp) owns one job; lane j of the group owns C
//     adjacent k-mer columns (M, B, K states and the read-scaled Gaussian of each in registers) and walks
//     the event rows one step behind lane j-1 — a systolic wavefront over the anti-diagonals.  The left
//     neighbour's three states travel by __shfl_up_sync (segment width W); nothing of the
//     (E+1) x 3(K+2) matrix the reference mallocs per call is ever stored.
//   * W == 32 only: jobs wider than 32*C columns are cut into strips chained back-to-back (lane 0 starts
//     strip s+1 while lane 31 is still finishing strip s); only the strip's right-edge column (3 floats
//     per row) goes through an L1-resident scratch line.  W < 32 classes hold single-strip jobs (K <= W*C).
//   * the 62.8 KB quantised log-sum table lives in shared memory (exact_math.cuh: 8 instructions per sum).
//   * every float operation is issued in the reference's order with explicit round-to-nearest
//     intrinsics (no FMA contraction), IEEE division included, so scores are bit-identical.
//   * persistent CTAs (one per SM); warps pull 32/W jobs at a time, longest first, from an atomic counter.
#pragma once
#include "nph_internal.cuh"
#include "exact_math.cuh"
#include <math_constants.h>

namespace nph_fwd {

// the pair operations of the row update, packed or per half according to the knobs above (same values either way)
__device__ __forceinline__ f32x2 P_lsum2(f32x2 a, f32x2 b, const LogsumTable tb)
{
#if NPH_PACKED_LSUM
    return lsum2(a, b, tb);
#else
    return pk2(lsum(lo2(a), lo2(b), tb), lsum(hi2(a), hi2(b), tb));
#endif
}
__device__ __forceinline__ f32x2 P_add2(f32x2 a, f32x2 b)
{
#if NPH_PACKED_ARITH
    return add2(a, b);
#else
    return pk2(__fadd_rn(lo2(a), lo2(b)), __fadd_rn(hi2(a), hi2(b)));
#endif
}
// Gaussian log-density of a pair: cc + (-0.5 * z) * z with z = (x - mu) / sigma (emissions.h:51-55); nz2 = packed -0.0 (see the kernel)
__device__ __forceinline__ f32x2 P_emission(f32x2 x2, f32x2 mu2, f32x2 nsd2, f32x2 cc2, f32x2 ry2, f32x2 nz2)
{
#if NPH_PACKED_ARITH
    const f32x2 a = div2_by_cached_rcp(sub2(x2, mu2), nsd2, ry2);
    return add2(cc2, fma2(mul2(bc2(-0.5f), a), a, nz2));
#else
    const float a0 = div_by_cached_rcp(__fsub_rn(lo2(x2), lo2(mu2)), -lo2(nsd2), lo2(ry2));
    const float a1 = div_by_cached_rcp(__fsub_rn(hi2(x2), hi2(mu2)), -hi2(nsd2), hi2(ry2));
    return pk2(__fadd_rn(lo2(cc2), __fmul_rn(__fmul_rn(-0.5f, a0), a0)), __fadd_rn(hi2(cc2), __fmul_rn(__fmul_rn(-0.5f, a1), a1)));
#endif
}

// Warps per persistent CTA (one CTA per SM).  The full-warp classes hold 16 warps at up to 128 registers; the sub-warp classes of short
// windows need fewer registers (80 at C = 4) and are latency bound on their per-step loads (issue-active 77 % with 16 warps,
// profiles/r02_hmm_forward_methylation_classes_summary.md), so they run 24 (C <= 4) or 20 (C <= 6) warps.
constexpr int kMaxWarpsPerCta = 24;
template <int C, int W> struct CtaShape { static constexpr int warps = (W == 32) ? 16 : (C <= 4 ? 24 : (C <= 6 ? 20 : 16)); };
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMinPeriod = 40;   // chained strips: the right edge of row r must be written >32 steps before it is read

struct FwdParams {
    const float* level;           // drift-scaled event levels, all reads
    const DevRead* reads;
    const float2* trans;          // per read (lp_mm_self, lp_mm_next)
    const DevModelView* models;
    const uint32_t* ranks;
    const nph_hmm_job* jobs;
    const uint32_t* order;        // this class's slice of the schedule
    uint32_t n_jobs;
    unsigned int* counter;
    const float* logsum_g;
    const float* flank;
    float* scores;
    float4* scratch_params;       // per warp: kpad_stride float4 {mu', sigma', log(1/sqrt(2pi)) - log sigma', RN(1/sigma')}
    float* scratch_edge;          // per warp: 3 * edge_stride floats (W == 32 classes)
    uint32_t kpad_stride;
    uint32_t edge_stride;
    uint32_t lsum_bias;           // NPH_LOGSUM_ADDR_BIAS, passed at run time on purpose (exact_math.cuh)
    uint32_t lsum_scale;          // 4, at run time for the same reason (keeps the table address an IMAD)
    float neg_zero;               // -0.0f, at run time: see the emission in the packed row update
    const uint32_t* progress;     // one-shot call: number of level chunks landed so far (nullptr: all resident)
    uint32_t chunk_events;        // events per level chunk (multiple of 32)
    HmmConsts c;
};

#ifndef NPH_PACKED_F32X2
#define NPH_PACKED_F32X2 1          // 0: scalar inner loop for every C (the round-1 kernel; kept for A/B measurements)
#endif
// A/B knobs of the paired row update (measurements in profiles/r02_k1_variants.md): which parts use packed instructions
#ifndef NPH_PACKED_LSUM
#define NPH_PACKED_LSUM 1           // 0: the log-sums of a pair as two scalar lsum
#endif
#ifndef NPH_PACKED_ARITH
#define NPH_PACKED_ARITH 1          // 0: transition adds and the Gaussian as scalar instructions on the halves
#endif

// CHAIN = false: every job of the class fits one strip (K <= W*C), all strip/edge bookkeeping compiles away.
// Even C: the lane's columns are paired (p, p + C/2) and the row update runs on sm_100's packed FP32 instructions
// (exact_math.cuh).  With that pairing the "left neighbour" operand of pair p is simply pair p-1, so no half ever has
// to be moved between register pairs except at the lane boundary; only the skip chain K[c] <- K[c-1], which is
// sequential across the columns of a row, stays scalar.
template <int C, int W, bool CHAIN>
__global__ void __launch_bounds__(CtaShape<C, W>::warps * 32, 1) hmm_forward_kernel(const FwdParams p)
{
    static_assert(W == 4 || W == 8 || W == 16 || W == 32, "group width");
    static_assert(!CHAIN || W == 32, "only full-warp groups chain strips");
    constexpr int G = 32 / W;                 // jobs per warp
    constexpr int kWarpsPerCta = CtaShape<C, W>::warps;
    constexpr int kCtaThreads = kWarpsPerCta * 32;
    constexpr int STRIP = W * C;              // columns per strip
    extern __shared__ float s_tbl[];
    for (int i = threadIdx.x; i < NPH_TBL_SMEM; i += kCtaThreads) s_tbl[i] = p.logsum_g[i];
    __syncthreads();
    const LogsumTable tb = make_logsum_table(s_tbl, p.lsum_bias, p.lsum_scale);

    const int lane = threadIdx.x & 31;
    const int gl = lane & (W - 1);            // lane within the group
    const int grp = lane / W;
    const int warp_global = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    float4* const my_params = p.scratch_params + (size_t)warp_global * p.kpad_stride + (W < 32 ? grp * STRIP : 0);
    float* const edge_m = p.scratch_edge + (size_t)warp_global * 3 * p.edge_stride;
    float* const edge_b = edge_m + p.edge_stride;
    float* const edge_k = edge_b + p.edge_stride;

    const float NEG = -CUDART_INF_F;
    const float lp_mk = p.c.lp_mk, lp_mb = p.c.lp_mb, lp_bb = p.c.lp_bb, lp_bk = p.c.lp_bk;
    const float lp_bm_next = p.c.lp_bm_next, lp_bm_self = p.c.lp_bm_self, lp_kk = p.c.lp_kk, lp_km = p.c.lp_km;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(p.counter, (unsigned)G);
        base = __shfl_sync(kFull, base, 0);
        if (base >= p.n_jobs) break;
        const uint32_t slot = base + grp;
        const bool has_job = slot < p.n_jobs;
        const uint32_t job_idx = p.order[has_job ? slot : base];
        const nph_hmm_job job = p.jobs[job_idx];
        const DevRead rd = p.reads[job.read];
        const float2 tr = p.trans[job.read];
        const float lp_mm_self = tr.x, lp_mm_next = tr.y;
        const DevModelView mv = p.models[job.model_id];
        if (p.progress) {
            // levels still streaming in behind us: wait until the chunk holding this read's last event has landed
            // (chunks land in order; the progress word is written by a copy queued behind the chunk's data)
            uint32_t need = has_job ? (uint32_t)((rd.event_off + rd.n_events - 1) / p.chunk_events) + 1u : 0u;
            need = __reduce_max_sync(kFull, need);
            if (lane == 0) {
                const volatile uint32_t* pr = p.progress;
                while (*pr < need) __nanosleep(500);
            }
            __syncwarp();
        }

        const int K = (int)job.n_kmers;
        const int E = has_job ? (int)(job.event_stop > job.event_start ? job.event_stop - job.event_start
                                                                       : job.event_start - job.event_stop) + 1
                              : 0;
        const int stride = job.stride;
        const bool pre_clip = (job.flags & NPH_HAF_ALLOW_PRE_CLIP) != 0;
        const bool post_clip = (job.flags & NPH_HAF_ALLOW_POST_CLIP) != 0;
        const int n_strips = CHAIN ? (K + STRIP - 1) / STRIP : 1;
        const int kpad = n_strips * STRIP;
        const int P = n_strips > 1 ? max(E, kMinPeriod) : E;

        // ---- per-job prologue: read-scaled Gaussian of every k-mer, formed in FP64 exactly as
        // get_scaled_gaussian_from_pore_model_state does, then narrowed; plus RN(1/sigma') ----
        if (has_job) {
            const uint32_t* rk = p.ranks + job.rank_off;
            for (int i = gl; i < kpad; i += W) {
                float4 g = make_float4(0.f, 1.f, 0.f, 1.f);
                if (i < K) {
                    const uint32_t r = rk[i];
                    const float mu = (float)__dadd_rn(__dmul_rn(rd.scale, mv.mean[r]), rd.shift);
                    const float sd = (float)__dmul_rn(mv.stdv[r], rd.var);
                    const float lsd = (float)__dadd_rn(mv.log_stdv[r], rd.log_var);
                    g = make_float4(mu, sd, __fsub_rn(p.c.log_inv_sqrt_2pi, lsd), __frcp_rn(sd));
                }
                my_params[i] = g;
            }
        }
        __syncwarp();

        const float* lv = p.level + rd.event_off;
        const long long e_first = (long long)job.event_start;

        // the lane/slot that owns the last k-mer column (end-state fold)
        const int last_strip = n_strips - 1;
        const int last_rel = (K - 1) - last_strip * STRIP;
        const int end_lane = last_rel / C;
        const int end_slot = last_rel % C;
        // lanes beyond end_lane own no column of the last strip; the warp runs until its slowest group is done
        const int my_steps = has_job ? last_strip * P + E + end_lane : 0;
        const int total_steps = (W == 32) ? my_steps : __reduce_max_sync(kFull, my_steps);

        constexpr bool PACKED = (NPH_PACKED_F32X2 != 0) && (C % 2 == 0);
        constexpr int H = PACKED ? C / 2 : 1;               // pairs per lane: pair p = columns (p, p + H)
        constexpr int CS = PACKED ? 1 : C;                  // scalar state arrays collapse to one unused slot when packed
        float mu[CS], sd[CS], cc[CS], ry[CS];
        float Mp[CS], Bp[CS], Kp[CS];
        f32x2 mu2[H], nsd2[H], cc2[H], ry2[H];              // Gaussian of the pair's columns; nsd = -sigma' (division by fma)
        f32x2 Mp2[H], Bp2[H], Kp2[H];
#pragma unroll
        for (int c = 0; c < CS; ++c) { mu[c] = 0.f; sd[c] = 1.f; cc[c] = 0.f; ry[c] = 1.f; Mp[c] = NEG; Bp[c] = NEG; Kp[c] = NEG; }
#pragma unroll
        for (int q = 0; q < H; ++q) {
            mu2[q] = bc2(0.f); nsd2[q] = bc2(-1.f); cc2[q] = bc2(0.f); ry2[q] = bc2(1.f);
            Mp2[q] = bc2(NEG); Bp2[q] = bc2(NEG); Kp2[q] = bc2(NEG);
        }
        float Lm_prev = NEG, Lb_prev = NEG, Lk_prev = NEG;
        float lp_end = NEG;
        int r = 1 - gl;        // row of this lane at the current step (rows 1..P; <1 = not started)
        int s = 0;             // strip of this lane
        float x_next = 0.f;
        if (r == 1 && E >= 1) x_next = lv[e_first];
        float em_next = NEG, eb_next = NEG, ek_next = NEG;   // group lane 0: prefetched right edge of the previous strip

        for (int g = 0; g < total_steps; ++g) {
            // left neighbour's newest row (its row == my row, computed one step ago)
            float Lm = __shfl_up_sync(kFull, PACKED ? hi2(Mp2[H - 1]) : Mp[CS - 1], 1, W);
            float Lb = __shfl_up_sync(kFull, PACKED ? hi2(Bp2[H - 1]) : Bp[CS - 1], 1, W);
            float Lk = __shfl_up_sync(kFull, PACKED ? hi2(Kp2[H - 1]) : Kp[CS - 1], 1, W);
            if (gl == 0) { Lm = CHAIN ? em_next : NEG; Lb = CHAIN ? eb_next : NEG; Lk = CHAIN ? ek_next : NEG; }

            const bool in_strip = (r >= 1) && (s < n_strips);
            const int col0 = s * STRIP + gl * C;
            const bool live = in_strip && (r <= E) && (col0 < K);
            const float x = x_next;

            if (in_strip && r == 1) {
                // entering a strip: row 0 and the start column are -inf
#pragma unroll
                for (int c = 0; c < CS; ++c) { Mp[c] = NEG; Bp[c] = NEG; Kp[c] = NEG; }
#pragma unroll
                for (int q = 0; q < H; ++q) { Mp2[q] = bc2(NEG); Bp2[q] = bc2(NEG); Kp2[q] = bc2(NEG); }
                Lm_prev = NEG; Lb_prev = NEG; Lk_prev = NEG;
                if (col0 < K) {
                    if (PACKED) {
#pragma unroll
                        for (int q = 0; q < H; ++q) {
                            const float4 ga = my_params[col0 + q], gb = my_params[col0 + q + H];
                            mu2[q] = pk2(ga.x, gb.x); nsd2[q] = pk2(-ga.y, -gb.y); cc2[q] = pk2(ga.z, gb.z); ry2[q] = pk2(ga.w, gb.w);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < CS; ++c) {
                            const float4 g4 = my_params[col0 + c];
                            mu[c] = g4.x; sd[c] = g4.y; cc[c] = g4.z; ry[c] = g4.w;
                        }
                    }
                }
            }

            // prefetch for the next step: event level and (group lane 0, chained strips) the stored right edge
            {
                int rn = r + 1, sn = s;
                if (rn > P) { rn = 1; sn = s + 1; }
                if (rn >= 1 && rn <= E && sn < n_strips) {
                    x_next = lv[e_first + (long long)(rn - 1) * stride];
                    if (CHAIN && gl == 0 && sn > 0) { em_next = edge_m[rn]; eb_next = edge_b[rn]; ek_next = edge_k[rn]; }
                }
            }

            if (live) {
                float soft = NEG;
                if (col0 == 0 && (r == 1 || pre_clip)) soft = p.flank[r - 1];
                float post = 0.f;
                const bool do_end = (s == last_strip) && (gl == end_lane) && (post_clip || r == E);
                if (do_end) post = p.flank[E - r];
                float Me, Be, Ke;                                                  // states of the last k-mer's column (do_end)

                if (PACKED) {
                    const f32x2 x2 = bc2(x);
                    // ptxas contracts mul.rn.f32x2 feeding add.rn.f32x2 into one FFMA2 (-fmad=false and the .rn modifiers do not
                    // stop it for the packed forms; seen in SASS, caught by the bit-parity tests).  The product is therefore
                    // formed as fma(t, a, -0.0) with a -0.0 ptxas cannot see (x*y + -0.0 == RN(x*y) including the sign of zero),
                    // which leaves FFMA2 -> FADD2, a pair that cannot be fused.
                    const f32x2 nz2 = bc2(p.neg_zero);
                    f32x2 mN[H], bN[H], xk[H];
#pragma unroll
                    for (int q = 0; q < H; ++q) {
                        // Gaussian log-density of both columns, reference operation order (emissions.h:51-55)
                        const f32x2 em = P_emission(x2, mu2[q], nsd2[q], cc2[q], ry2[q], nz2);
                        // left column, previous row: pair q-1 as it stands; at the lane boundary the neighbour's value and column H-1
                        const f32x2 sM = q ? Mp2[q > 0 ? q - 1 : 0] : pk2(Lm_prev, lo2(Mp2[H - 1]));
                        const f32x2 sB = q ? Bp2[q > 0 ? q - 1 : 0] : pk2(Lb_prev, lo2(Bp2[H - 1]));
                        const f32x2 sK = q ? Kp2[q > 0 ? q - 1 : 0] : pk2(Lk_prev, lo2(Kp2[H - 1]));
                        // match: left fold over {same M, prev M, same B, prev B, prev K, soft}
                        f32x2 m = P_add2(bc2(lp_mm_self), Mp2[q]);
                        m = P_lsum2(m, P_add2(bc2(lp_mm_next), sM), tb);
                        m = P_lsum2(m, P_add2(bc2(lp_bm_self), Bp2[q]), tb);
                        m = P_lsum2(m, P_add2(bc2(lp_bm_next), sB), tb);
                        m = P_lsum2(m, P_add2(bc2(lp_km), sK), tb);
                        if (q == 0) m = pk2(lsum(lo2(m), soft, tb), hi2(m));       // column 0 of the lane only
                        mN[q] = P_add2(m, em);
                        // bad event: {same M, same B}
                        bN[q] = P_lsum2(P_add2(bc2(lp_mb), Mp2[q]), P_add2(bc2(lp_bb), Bp2[q]), tb);
                    }
#pragma unroll
                    for (int q = 0; q < H; ++q) {
                        // k-mer skip, the part that does not depend on the chain: {prev M, prev B} of the SAME row
                        const f32x2 cM = q ? mN[q > 0 ? q - 1 : 0] : pk2(Lm, lo2(mN[H - 1]));
                        const f32x2 cB = q ? bN[q > 0 ? q - 1 : 0] : pk2(Lb, lo2(bN[H - 1]));
                        xk[q] = P_lsum2(P_add2(bc2(lp_mk), cM), P_add2(bc2(lp_bk), cB), tb);
                    }
                    // the chain K[c] = x[c] (+) (lp_kk + K[c-1]) runs through the columns in order: lo halves, then hi halves
                    float kn[C];
                    float kprev = Lk;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float xc = c < H ? lo2(xk[c < H ? c : 0]) : hi2(xk[c >= H ? c - H : 0]);
                        kn[c] = lsum(xc, __fadd_rn(lp_kk, kprev), tb);
                        kprev = kn[c];
                    }
#pragma unroll
                    for (int q = 0; q < H; ++q) { Mp2[q] = mN[q]; Bp2[q] = bN[q]; Kp2[q] = pk2(kn[q], kn[q + H]); }
                    Me = lo2(Mp2[0]); Be = lo2(Bp2[0]); Ke = lo2(Kp2[0]);
#pragma unroll
                    for (int c = 1; c < C; ++c)
                        if (c == end_slot) {
                            Me = c < H ? lo2(Mp2[c < H ? c : 0]) : hi2(Mp2[c >= H ? c - H : 0]);
                            Be = c < H ? lo2(Bp2[c < H ? c : 0]) : hi2(Bp2[c >= H ? c - H : 0]);
                            Ke = c < H ? lo2(Kp2[c < H ? c : 0]) : hi2(Kp2[c >= H ? c - H : 0]);
                        }
                } else {
                    float lm_prev = Lm_prev, lb_prev = Lb_prev, lk_prev = Lk_prev;   // left column, row r-1
                    float lm_cur = Lm, lb_cur = Lb, lk_cur = Lk;                      // left column, row r
#pragma unroll
                    for (int c = 0; c < CS; ++c) {
                        // Gaussian log-density, reference operation order (emissions.h:51-55)
                        const float a = div_by_cached_rcp(__fsub_rn(x, mu[c]), sd[c], ry[c]);
                        const float em = __fadd_rn(cc[c], __fmul_rn(__fmul_rn(-0.5f, a), a));
                        // match: left fold over {same M, prev M, same B, prev B, prev K, soft}
                        float m = __fadd_rn(lp_mm_self, Mp[c]);
                        m = lsum(m, __fadd_rn(lp_mm_next, lm_prev), tb);
                        m = lsum(m, __fadd_rn(lp_bm_self, Bp[c]), tb);
                        m = lsum(m, __fadd_rn(lp_bm_next, lb_prev), tb);
                        m = lsum(m, __fadd_rn(lp_km, lk_prev), tb);
                        if (c == 0) m = lsum(m, soft, tb);
                        m = __fadd_rn(m, em);
                        // bad event: {same M, same B}
                        const float b = lsum(__fadd_rn(lp_mb, Mp[c]), __fadd_rn(lp_bb, Bp[c]), tb);
                        // k-mer skip: {prev M, prev B, prev K} of the SAME row
                        float kk = lsum(__fadd_rn(lp_mk, lm_cur), __fadd_rn(lp_bk, lb_cur), tb);
                        kk = lsum(kk, __fadd_rn(lp_kk, lk_cur), tb);

                        lm_prev = Mp[c]; lb_prev = Bp[c]; lk_prev = Kp[c];
                        lm_cur = m; lb_cur = b; lk_cur = kk;
                        Mp[c] = m; Bp[c] = b; Kp[c] = kk;
                    }
                    Me = Mp[0]; Be = Bp[0]; Ke = Kp[0];
#pragma unroll
                    for (int c = 1; c < CS; ++c) if (c == end_slot) { Me = Mp[c]; Be = Bp[c]; Ke = Kp[c]; }
                }
                Lm_prev = Lm; Lb_prev = Lb; Lk_prev = Lk;

                if (do_end) {
                    lp_end = lsum(lp_end, __fadd_rn(Me, post), tb);
                    lp_end = lsum(lp_end, __fadd_rn(Be, post), tb);
                    lp_end = lsum(lp_end, __fadd_rn(Ke, post), tb);
                }
                if (CHAIN && gl == W - 1 && s < last_strip) {
                    edge_m[r] = PACKED ? hi2(Mp2[H - 1]) : Mp[CS - 1];
                    edge_b[r] = PACKED ? hi2(Bp2[H - 1]) : Bp[CS - 1];
                    edge_k[r] = PACKED ? hi2(Kp2[H - 1]) : Kp[CS - 1];
                }
            }

            // advance
            r += 1;
            if (r > P) { r = 1; s += 1; }
            if (CHAIN && n_strips > 1) __syncwarp();   // orders lane 31's edge stores before lane 0's later loads
        }

        const float result = __shfl_sync(kFull, lp_end, (lane & ~(W - 1)) + end_lane);
        if (gl == 0 && has_job) p.scores[job_idx] = result;
        __syncwarp();
    }
}

template <int C, int W, bool CHAIN>
int launch_class(nph_ctx* ctx, const FwdParams& base, const nph_ctx::ClassLaunch& cl, int class_idx, cudaStream_t stream)
{
    FwdParams p = base;
    p.order = ctx->d_order.p + cl.first;
    p.n_jobs = (uint32_t)cl.count;
    p.counter = ctx->d_counters.p + class_idx;
    const size_t smem = sizeof(float) * NPH_TBL_SMEM;
    NPH_CUDA(ctx, cudaFuncSetAttribute(hmm_forward_kernel<C, W, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    constexpr int kWarpsPerCta = CtaShape<C, W>::warps;
    int grid = ctx->sm_count;
    const size_t warps_needed = (cl.count + (32 / W) - 1) / (32 / W);
    if ((size_t)grid * kWarpsPerCta > warps_needed) grid = (int)((warps_needed + kWarpsPerCta - 1) / kWarpsPerCta);
    if (grid < 1) grid = 1;
    hmm_forward_kernel<C, W, CHAIN><<<grid, kWarpsPerCta * 32, smem, stream>>>(p);
    NPH_CUDA(ctx, cudaGetLastError());
    return NPH_OK;
}

// one translation unit per group width (parallel compilation); defined in hmm_forward_w*.cu
template <int W, bool CHAIN> int launch_width(nph_ctx* ctx, const FwdParams& base, const nph_ctx::ClassLaunch& cl, int class_idx, cudaStream_t stream);

#define NPH_DEFINE_LAUNCH_WIDTH(W, CHAIN)                                                                              \
    template <> int launch_width<W, CHAIN>(nph_ctx * ctx, const FwdParams& base, const nph_ctx::ClassLaunch& cl, int class_idx, cudaStream_t stream) \
    {                                                                                                           \
        switch (cl.cols_per_lane) {                                                                             \
            case 1: return launch_class<1, W, CHAIN>(ctx, base, cl, class_idx, stream);                                       \
            case 2: return launch_class<2, W, CHAIN>(ctx, base, cl, class_idx, stream);                                       \
            case 3: return launch_class<3, W, CHAIN>(ctx, base, cl, class_idx, stream);                                       \
            case 4: return launch_class<4, W, CHAIN>(ctx, base, cl, class_idx, stream);                                       \
            case 5: return launch_class<5, W, CHAIN>(ctx, base, cl, class_idx, stream);                                       \
            case 6: return launch_class<6, W, CHAIN>(ctx, base, cl, class_idx, stream);                                       \
            case 7: return launch_class<7, W, CHAIN>(ctx, base, cl, class_idx, stream);                                       \
            case 8: return launch_class<8, W, CHAIN>(ctx, base, cl, class_idx, stream);                                       \
            case 9: return launch_class<9, W, CHAIN>(ctx, base, cl, class_idx, stream);                                       \
            case 10: return launch_class<10, W, CHAIN>(ctx, base, cl, class_idx, stream);                                     \
        }                                                                                                       \
        return NPH_ERR_STATE;                                                                                   \
    }

} // namespace nph_fwd
