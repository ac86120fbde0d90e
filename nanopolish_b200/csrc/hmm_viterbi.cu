// hmm_viterbi.cu — K3: Viterbi alignment of events to the k-mers of a sequence.
//
// Replaces, for a batch of (sequence, HMMInputData, flags) jobs:
//   profile_hmm_align_r9          ref: src/hmm/nanopolish_profile_hmm_r9.cpp:73-204
//   ProfileHMMViterbiOutputR9     ref: src/hmm/nanopolish_profile_hmm_r9.inl:130-197
//   (the fill itself is profile_hmm_fill_generic_r9, .inl:265-433, shared with the forward score)
//
// Same systolic mapping as the forward kernel (hmm_forward_kernel.cuh): a warp per job, lane j owns C
// k-mer columns, strips of 32*C columns chained.  Differences:
//   * (+) is max with the reference's argmax rule — a compare chain over the six movement types in
//     index order where a LATER index wins ties (`from = max == x[i] ? i : from`), so an all -inf cell
//     records FROM_SOFT like the reference does;
//   * each block-cell stores its three 3-bit movement codes (9 bits in a uint16) to a per-warp trace
//     laid out by systolic step, so the 32 lanes of a step write one contiguous line;
//   * the reference keeps the whole float matrix only to report l_fm along the path; we keep no
//     values: after the backtrack the path is replayed forwards, and l_fm of each state is recomputed
//     from its predecessor with the very float operations the fill used (transition + value, + emission),
//     which reproduces the stored matrix entries bit for bit;
//   * the backtrack starts at (last event, MATCH of the last k-mer) and stops at FROM_SOFT, as the
//     reference does; where the reference would trip an assert (fewer than 2 events, path entering a
//     -inf cell) the job returns zero states.
#include "nph_internal.cuh"
#include "exact_math.cuh"
#include <math_constants.h>
#include <algorithm>
#include <vector>

#define NPH_TRY(expr) do { int rc__ = (expr); if (rc__ != NPH_OK) return rc__; } while (0)

namespace {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMinPeriod = 40;
enum { MV_SAME_M = 0, MV_PREV_M = 1, MV_SAME_B = 2, MV_PREV_B = 3, MV_PREV_K = 4, MV_SOFT = 5 };

struct VitParams {
    const float* level;
    const DevRead* reads;
    const float2* trans;
    const DevModelView* models;
    const uint32_t* ranks;
    const nph_hmm_job* jobs;
    const uint64_t* states_off;      // n_jobs + 1
    const uint32_t* order;
    uint32_t n_jobs;
    unsigned int* counter;
    const float* flank;
    nph_align_state* states;
    uint32_t* n_states;
    float* scores;
    float4* scratch_params;
    float* scratch_edge;
    uint16_t* scratch_trace;
    uint32_t kpad_stride, edge_stride;
    uint64_t trace_stride;           // uint16 elements per warp
    HmmConsts c;
};

// running max with the reference's tie rule: a later candidate that equals the max takes the label
__device__ __forceinline__ void vmax(float& mx, int& from, float x, int idx)
{
    mx = x > mx ? x : mx;
    from = (mx == x) ? idx : from;
}

template <int C>
__global__ void __launch_bounds__(kThreads, 1) hmm_viterbi_kernel(const VitParams p)
{
    constexpr int STRIP = 32 * C;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kWarps + (threadIdx.x >> 5);
    float4* const my_params = p.scratch_params + (size_t)warp_global * p.kpad_stride;
    float* const edge_m = p.scratch_edge + (size_t)warp_global * 3 * p.edge_stride;
    float* const edge_b = edge_m + p.edge_stride;
    float* const edge_k = edge_b + p.edge_stride;
    uint16_t* const trace = p.scratch_trace + (size_t)warp_global * p.trace_stride;
    const float NEG = -CUDART_INF_F;
    const float lp_mk = p.c.lp_mk, lp_mb = p.c.lp_mb, lp_bb = p.c.lp_bb, lp_bk = p.c.lp_bk;
    const float lp_bm_next = p.c.lp_bm_next, lp_bm_self = p.c.lp_bm_self, lp_kk = p.c.lp_kk, lp_km = p.c.lp_km;

    for (;;) {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(p.counter, 1u);
        slot = __shfl_sync(kFull, slot, 0);
        if (slot >= p.n_jobs) break;
        const uint32_t job_idx = p.order[slot];
        const nph_hmm_job job = p.jobs[job_idx];
        const DevRead rd = p.reads[job.read];
        const float2 tr = p.trans[job.read];
        const float lp_mm_self = tr.x, lp_mm_next = tr.y;
        const DevModelView mv = p.models[job.model_id];
        const int K = (int)job.n_kmers;
        const int E = (int)(job.event_stop > job.event_start ? job.event_stop - job.event_start : job.event_start - job.event_stop) + 1;
        const int stride = job.stride;
        const bool pre_clip = (job.flags & NPH_HAF_ALLOW_PRE_CLIP) != 0;
        const int n_strips = (K + STRIP - 1) / STRIP;
        const int kpad = n_strips * STRIP;
        const int P = n_strips > 1 ? max(E, kMinPeriod) : E;
        nph_align_state* const out = p.states + p.states_off[job_idx];
        const int cap = (int)(p.states_off[job_idx + 1] - p.states_off[job_idx]);

        if (E < 2) {                     // the reference asserts n_events >= 2 (profile_hmm_r9.cpp:88)
            if (lane == 0) { p.n_states[job_idx] = 0; if (p.scores) p.scores[job_idx] = NEG; }
            continue;
        }
        {
            const uint32_t* rk = p.ranks + job.rank_off;
            for (int i = lane; i < kpad; i += 32) {
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
        const int last_strip = n_strips - 1;
        const int end_lane = ((K - 1) - last_strip * STRIP) / C;
        const int total_steps = last_strip * P + E + end_lane;

        // ---------------------------------- fill ----------------------------------
        float mu[C], sd[C], cc[C], ry[C], Mp[C], Bp[C], Kp[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { mu[c] = 0.f; sd[c] = 1.f; cc[c] = 0.f; ry[c] = 1.f; Mp[c] = NEG; Bp[c] = NEG; Kp[c] = NEG; }
        float Lm_prev = NEG, Lb_prev = NEG, Lk_prev = NEG;
        int r = 1 - lane, s = 0;
        float x_next = 0.f;
        if (r == 1) x_next = lv[e_first];
        float em_next = NEG, eb_next = NEG, ek_next = NEG;

        for (int g = 0; g < total_steps; ++g) {
            float Lm = __shfl_up_sync(kFull, Mp[C - 1], 1);
            float Lb = __shfl_up_sync(kFull, Bp[C - 1], 1);
            float Lk = __shfl_up_sync(kFull, Kp[C - 1], 1);
            if (lane == 0) { Lm = em_next; Lb = eb_next; Lk = ek_next; }
            const bool in_strip = (r >= 1) && (s < n_strips);
            const int col0 = s * STRIP + lane * C;
            const bool live = in_strip && (r <= E) && (col0 < K);
            const float x = x_next;
            if (in_strip && r == 1) {
#pragma unroll
                for (int c = 0; c < C; ++c) { Mp[c] = NEG; Bp[c] = NEG; Kp[c] = NEG; }
                Lm_prev = NEG; Lb_prev = NEG; Lk_prev = NEG;
                if (col0 < K) {
#pragma unroll
                    for (int c = 0; c < C; ++c) { const float4 g4 = my_params[col0 + c]; mu[c] = g4.x; sd[c] = g4.y; cc[c] = g4.z; ry[c] = g4.w; }
                }
            }
            {
                int rn = r + 1, sn = s;
                if (rn > P) { rn = 1; sn = s + 1; }
                if (rn >= 1 && rn <= E && sn < n_strips) {
                    x_next = lv[e_first + (long long)(rn - 1) * stride];
                    if (lane == 0 && sn > 0) { em_next = edge_m[rn]; eb_next = edge_b[rn]; ek_next = edge_k[rn]; }
                }
            }
            uint16_t tcode[C];
#pragma unroll
            for (int c = 0; c < C; ++c) tcode[c] = 0;
            if (live) {
                float soft = NEG;
                if (col0 == 0 && (r == 1 || pre_clip)) soft = __fadd_rn(0.0f, p.flank[r - 1]);
                float lm_prev = Lm_prev, lb_prev = Lb_prev, lk_prev = Lk_prev;
                float lm_cur = Lm, lb_cur = Lb, lk_cur = Lk;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float a = div_by_cached_rcp(__fsub_rn(x, mu[c]), sd[c], ry[c]);
                    const float em = __fadd_rn(cc[c], __fmul_rn(__fmul_rn(-0.5f, a), a));
                    // MATCH: six candidates in movement order
                    float m = __fadd_rn(lp_mm_self, Mp[c]);
                    int fm = MV_SAME_M;
                    vmax(m, fm, __fadd_rn(lp_mm_next, lm_prev), MV_PREV_M);
                    vmax(m, fm, __fadd_rn(lp_bm_self, Bp[c]), MV_SAME_B);
                    vmax(m, fm, __fadd_rn(lp_bm_next, lb_prev), MV_PREV_B);
                    vmax(m, fm, __fadd_rn(lp_km, lk_prev), MV_PREV_K);
                    vmax(m, fm, (c == 0) ? soft : NEG, MV_SOFT);
                    m = __fadd_rn(m, em);
                    // BAD EVENT: {same M, -inf, same B, -inf, -inf, -inf}
                    float b = __fadd_rn(lp_mb, Mp[c]);
                    int fb = MV_SAME_M;
                    fb = (b == NEG) ? MV_PREV_M : fb;
                    vmax(b, fb, __fadd_rn(lp_bb, Bp[c]), MV_SAME_B);
                    fb = (b == NEG) ? MV_SOFT : fb;
                    // K-MER SKIP: {-inf, prev M, -inf, prev B, prev K, -inf} of the same row
                    float kk = __fadd_rn(lp_mk, lm_cur);
                    int fk = MV_PREV_M;
                    fk = (kk == NEG) ? MV_SAME_B : fk;
                    vmax(kk, fk, __fadd_rn(lp_bk, lb_cur), MV_PREV_B);
                    vmax(kk, fk, __fadd_rn(lp_kk, lk_cur), MV_PREV_K);
                    fk = (kk == NEG) ? MV_SOFT : fk;

                    lm_prev = Mp[c]; lb_prev = Bp[c]; lk_prev = Kp[c];
                    lm_cur = m; lb_cur = b; lk_cur = kk;
                    Mp[c] = m; Bp[c] = b; Kp[c] = kk;
                    tcode[c] = (uint16_t)(fm | (fb << 3) | (fk << 6));
                }
                Lm_prev = Lm; Lb_prev = Lb; Lk_prev = Lk;
                if (lane == 31 && s < last_strip) { edge_m[r] = Mp[C - 1]; edge_b[r] = Bp[C - 1]; edge_k[r] = Kp[C - 1]; }
            }
            // trace line of this step: one contiguous 64*C bytes per warp
#pragma unroll
            for (int c = 0; c < C; ++c) trace[(size_t)g * STRIP + lane * C + c] = tcode[c];
            r += 1;
            if (r > P) { r = 1; s += 1; }
            if (n_strips > 1) __syncwarp();
        }
        __syncwarp();

        // ---------------------------------- backtrack (all lanes walk the same path) ----------------------------------
        int n = 0, status = 0;
        {
            int row = E, kmer = K - 1, st = 2;        // state codes: 0 K, 1 B, 2 M (column % 3 in the reference)
            while (row > 0) {
                const int sidx = kmer / STRIP, rel = kmer - sidx * STRIP;
                const int step = sidx * P + (row - 1) + rel / C;
                const uint32_t code = __ldcg(trace + (size_t)step * STRIP + rel);
                const int mvt = (st == 2) ? (code & 7) : (st == 1) ? ((code >> 3) & 7) : ((code >> 6) & 7);
                if (n >= cap) { status = 3; break; }
                if (lane == 0) {
                    nph_align_state a;
                    a.event_idx = (uint32_t)(e_first + (long long)(row - 1) * stride);
                    a.kmer_idx = (uint32_t)kmer;
                    a.l_fm = 0.f;
                    a.state = (st == 2) ? 'M' : (st == 1) ? 'B' : 'K';
                    a.reserved[0] = (uint8_t)mvt; a.reserved[1] = 0; a.reserved[2] = 0;
                    out[cap - 1 - n] = a;
                }
                ++n;
                if (mvt == MV_SOFT) break;
                int nst = 2;
                switch (mvt) {
                    case MV_SAME_M: nst = 2; break;
                    case MV_PREV_M: kmer -= 1; nst = 2; break;
                    case MV_SAME_B: nst = 1; break;
                    case MV_PREV_B: kmer -= 1; nst = 1; break;
                    case MV_PREV_K: kmer -= 1; nst = 0; break;
                }
                if (st != 0) row -= 1;               // a k-mer skip is silent
                st = nst;
                if (kmer < 0) { status = 2; break; } // block 0: the reference asserts
            }
        }
        __syncwarp();

        // ---------------------------------- replay forwards: l_fm of every state ----------------------------------
        float last_v = NEG;
        if (!status && lane == 0) {
            float v = NEG;
            for (int i = 0; i < n; ++i) {
                nph_align_state a = out[cap - n + i];
                const int mvt = a.reserved[0];
                const int row = (int)(((long long)a.event_idx - e_first) * stride) + 1;
                float x5 = NEG;
                if (mvt == MV_SOFT) {
                    // legitimate only as the first state: MATCH of k-mer 0 at row 1 or with PRE_CLIP; anything else is a -inf cell
                    if (i == 0 && a.state == 'M' && a.kmer_idx == 0 && (row == 1 || pre_clip)) x5 = __fadd_rn(0.0f, p.flank[row - 1]);
                    else { status = 2; break; }
                }
                float t;
                if (a.state == 'M') {
                    const float tr_ = mvt == MV_SAME_M ? lp_mm_self : mvt == MV_PREV_M ? lp_mm_next : mvt == MV_SAME_B ? lp_bm_self
                                      : mvt == MV_PREV_B ? lp_bm_next : lp_km;
                    t = (mvt == MV_SOFT) ? x5 : __fadd_rn(tr_, v);
                    const float4 g4 = my_params[a.kmer_idx];
                    const float aa = div_by_cached_rcp(__fsub_rn(lv[a.event_idx], g4.x), g4.y, g4.w);
                    t = __fadd_rn(t, __fadd_rn(g4.z, __fmul_rn(__fmul_rn(-0.5f, aa), aa)));
                } else if (a.state == 'B') {
                    t = __fadd_rn(mvt == MV_SAME_M ? lp_mb : lp_bb, v);
                } else {
                    t = __fadd_rn(mvt == MV_PREV_M ? lp_mk : mvt == MV_PREV_B ? lp_bk : lp_kk, v);
                }
                if (t == NEG) { status = 2; break; }  // the reference asserts vm != -inf on every visited cell
                v = t;
                a.l_fm = v;
                a.reserved[0] = 0;
                out[i] = a;                            // compaction to the front: i <= cap - n + i, read before write
            }
            last_v = v;
        }
        status = __shfl_sync(kFull, status, 0);
        if (lane == 0) {
            p.n_states[job_idx] = status ? 0u : (uint32_t)n;
            if (p.scores) p.scores[job_idx] = status ? NEG : last_v;
        }
        __syncwarp();
    }
}

template <int C>
int launch_vit(nph_ctx* ctx, const VitParams& base, const uint32_t* order, size_t count, unsigned int* counter)
{
    VitParams p = base;
    p.order = order; p.n_jobs = (uint32_t)count; p.counter = counter;
    int grid = ctx->sm_count;
    if ((size_t)grid * kWarps > count) grid = (int)((count + kWarps - 1) / kWarps);
    if (grid < 1) grid = 1;
    hmm_viterbi_kernel<C><<<grid, kThreads, 0, ctx->stream>>>(p);
    NPH_CUDA(ctx, cudaGetLastError());
    return NPH_OK;
}

const int kVitCols[] = {1, 2, 3, 4, 6, 8};
const int kNumVit = 6;

inline uint32_t vit_steps(uint32_t K, uint32_t E, int C)
{
    const uint32_t strip = 32u * C, n_strips = (K + strip - 1) / strip;
    const uint32_t P = n_strips > 1 ? std::max<uint32_t>(E, kMinPeriod) : E;
    return (n_strips - 1) * P + E + ((K - (n_strips - 1) * strip) - 1) / C;
}

} // namespace

extern "C" int nph_hmm_align_batch(nph_ctx* ctx,
                                   const nph_read* reads, size_t n_reads,
                                   const float* ev_mean, const double* ev_start_time, size_t n_events_total,
                                   const uint32_t* kmer_ranks, size_t n_ranks_total,
                                   const nph_hmm_job* jobs, size_t n_jobs, double indel_bias,
                                   nph_align_state* states_out, const uint64_t* states_off,
                                   uint32_t* n_states_out, float* scores_out)
{
    if (!ctx || !jobs || !states_out || !states_off || !n_states_out || n_jobs == 0) return NPH_ERR_INVALID;
    NPH_TRY(nph_reads_load(ctx, reads, n_reads, ev_mean, ev_start_time, n_events_total));
    return nph_hmm_align(ctx, kmer_ranks, n_ranks_total, jobs, n_jobs, indel_bias, states_out, states_off, n_states_out, scores_out);
}

extern "C" int nph_hmm_align(nph_ctx* ctx,
                             const uint32_t* kmer_ranks, size_t n_ranks_total,
                             const nph_hmm_job* jobs, size_t n_jobs, double indel_bias,
                             nph_align_state* states_out, const uint64_t* states_off,
                             uint32_t* n_states_out, float* scores_out)
{
    if (!ctx || !jobs || !states_out || !states_off || !n_states_out || n_jobs == 0) return NPH_ERR_INVALID;
    if (!ctx->reads_loaded) return NPH_ERR_STATE;
    // jobs, ranks, transitions and validation go through the forward path's loader (same job semantics)
    NPH_TRY(nph_hmm_jobs_load(ctx, kmer_ranks, n_ranks_total, jobs, n_jobs, indel_bias));

    // class per job (columns per lane) and schedule, longest first
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> per(kNumVit);
    uint32_t max_kpad = 32, max_period = kMinPeriod;
    uint64_t max_trace = 1;
    for (size_t j = 0; j < n_jobs; ++j) {
        const nph_hmm_job& jb = jobs[j];
        const uint32_t E = (jb.event_stop > jb.event_start ? jb.event_stop - jb.event_start : jb.event_start - jb.event_stop) + 1;
        const uint32_t K = jb.n_kmers;
        if (states_off[j + 1] < states_off[j]) return NPH_ERR_INVALID;
        double best = 1e300; int bi = 0; uint32_t bsteps = 0;
        for (int i = 0; i < kNumVit; ++i) {
            const uint32_t st = vit_steps(K, E, kVitCols[i]);
            const double cost = (double)st * (120.0 + 70.0 * kVitCols[i]);
            if (cost < best) { best = cost; bi = i; bsteps = st; }
        }
        per[bi].push_back({bsteps, (uint32_t)j});
        const uint32_t strip = 32u * kVitCols[bi], n_strips = (K + strip - 1) / strip;
        max_kpad = std::max(max_kpad, n_strips * strip);
        max_period = std::max(max_period, std::max<uint32_t>(E, kMinPeriod));
        max_trace = std::max<uint64_t>(max_trace, (uint64_t)(bsteps + 1) * strip);
    }
    std::vector<uint32_t> order;
    order.reserve(n_jobs);
    size_t first[kNumVit + 1];
    for (int i = 0; i < kNumVit; ++i) {
        first[i] = order.size();
        std::sort(per[i].begin(), per[i].end(), [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) {
            return a.first != b.first ? a.first > b.first : a.second < b.second; });
        for (auto& e : per[i]) order.push_back(e.second);
    }
    first[kNumVit] = order.size();

    const int warps = ctx->sm_count * kWarps;
    const size_t total_states = (size_t)states_off[n_jobs];
    const size_t b_params = sizeof(float4) * (size_t)max_kpad * warps;
    const size_t b_edge = sizeof(float) * 3 * ((size_t)max_period + 8) * warps;
    const size_t trace_stride = ((max_trace + 63) / 64) * 64;
    const size_t b_trace = sizeof(uint16_t) * trace_stride * warps;
    const size_t b_states = sizeof(nph_align_state) * total_states;
    const size_t b_off = sizeof(uint64_t) * (n_jobs + 1);
    const size_t b_n = sizeof(uint32_t) * n_jobs;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t need = al(b_params) + al(b_edge) + al(b_trace) + al(b_states) + al(b_off) + al(b_n) + al(sizeof(uint32_t) * n_jobs);
    NPH_TRY(nph_reserve(ctx, ctx->d_abea_scratch, need));      // shares the alignment scratch arena with ABEA
    uint8_t* base = ctx->d_abea_scratch.p;
    VitParams p{};
    p.scratch_params = reinterpret_cast<float4*>(base); base += al(b_params);
    p.scratch_edge = reinterpret_cast<float*>(base); base += al(b_edge);
    p.scratch_trace = reinterpret_cast<uint16_t*>(base); base += al(b_trace);
    p.states = reinterpret_cast<nph_align_state*>(base); base += al(b_states);
    uint64_t* d_off = reinterpret_cast<uint64_t*>(base); base += al(b_off);
    p.n_states = reinterpret_cast<uint32_t*>(base); base += al(b_n);
    uint32_t* d_order = reinterpret_cast<uint32_t*>(base);
    p.states_off = d_off;
    p.level = ctx->d_level.p; p.reads = ctx->d_reads.p; p.trans = ctx->d_trans.p; p.models = ctx->d_models.p;
    p.ranks = ctx->d_ranks.p; p.jobs = ctx->d_jobs.p; p.flank = ctx->d_flank.p; p.scores = ctx->d_scores.p;
    p.kpad_stride = max_kpad; p.edge_stride = max_period + 8; p.trace_stride = trace_stride; p.c = ctx->consts;
    NPH_CUDA(ctx, cudaMemcpyAsync(d_off, states_off, b_off, cudaMemcpyHostToDevice, ctx->stream));
    NPH_CUDA(ctx, cudaMemcpyAsync(d_order, order.data(), sizeof(uint32_t) * n_jobs, cudaMemcpyHostToDevice, ctx->stream));
    NPH_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(unsigned int) * NPH_NUM_COUNTERS, ctx->stream));
    NPH_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    int launches = 0;
    for (int i = 0; i < kNumVit; ++i) {
        const size_t count = first[i + 1] - first[i];
        if (!count) continue;
        int rc = NPH_ERR_STATE;
        switch (kVitCols[i]) {
            case 1: rc = launch_vit<1>(ctx, p, d_order + first[i], count, ctx->d_counters.p + i); break;
            case 2: rc = launch_vit<2>(ctx, p, d_order + first[i], count, ctx->d_counters.p + i); break;
            case 3: rc = launch_vit<3>(ctx, p, d_order + first[i], count, ctx->d_counters.p + i); break;
            case 4: rc = launch_vit<4>(ctx, p, d_order + first[i], count, ctx->d_counters.p + i); break;
            case 6: rc = launch_vit<6>(ctx, p, d_order + first[i], count, ctx->d_counters.p + i); break;
            case 8: rc = launch_vit<8>(ctx, p, d_order + first[i], count, ctx->d_counters.p + i); break;
        }
        if (rc != NPH_OK) return rc;
        ++launches;
    }
    NPH_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->last_launches = launches;
    ctx->timing_valid = true;
    NPH_CUDA(ctx, cudaMemcpyAsync(states_out, p.states, b_states, cudaMemcpyDeviceToHost, ctx->stream));
    NPH_CUDA(ctx, cudaMemcpyAsync(n_states_out, p.n_states, b_n, cudaMemcpyDeviceToHost, ctx->stream));
    if (scores_out) NPH_CUDA(ctx, cudaMemcpyAsync(scores_out, ctx->d_scores.p, sizeof(float) * n_jobs, cudaMemcpyDeviceToHost, ctx->stream));
    NPH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->abea_loaded = false;   // the arena was reused
    return NPH_OK;
}
