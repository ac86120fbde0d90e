// hmm_classes.h — kernel-class choice for forward-HMM jobs, shared by host and device code.
//
// A class is (C columns per lane, W lanes per job); 32/W jobs share a warp.  W < 32 classes hold
// single-strip jobs (K <= W*C); W == 32 also chains strips for wide jobs.  A job goes to the class that
// minimises modelled issue slots = steps x (per-step overhead + C x per-cell cost) x W/32, with the
// constants measured by ncu (profiles/): ~96 per warp step, ~88 instructions per block-cell for odd C and ~66 for
// even C (whose columns run pairwise on sm_100's packed FP32 instructions).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NPH_HD __host__ __device__ __forceinline__
#else
#define NPH_HD inline
#endif

#define NPH_NUM_WIDTHS 5              // 4, 8, 16, 32 lanes single-strip; 32 lanes with chained strips
#define NPH_MAX_COLS 10
#define NPH_NUM_CLASSES (NPH_NUM_WIDTHS * NPH_MAX_COLS)
#define NPH_STEP_BUCKETS 1024         // step key: exact below 768, then 32-step bins
#define NPH_CHUNK_BUCKETS 8           // level chunk of the job's read (one-shot call), major key
#define NPH_KEY_BUCKETS (NPH_STEP_BUCKETS * NPH_CHUNK_BUCKETS)

NPH_HD uint32_t nph_class_width(int wi) { return wi >= 3 ? 32u : (4u << wi); }   // 4, 8, 16, 32, 32 (chained)
NPH_HD bool nph_class_chained(int wi) { return wi == 4; }
NPH_HD int nph_class_index(int C, int wi) { return wi * NPH_MAX_COLS + (C - 1); }

// warp steps one job takes in class (C, W): chained strips of W*C columns, period max(E, 40) when chained
NPH_HD uint32_t nph_class_steps(uint32_t K, uint32_t E, int C, uint32_t W)
{
    const uint32_t strip = W * (uint32_t)C;
    const uint32_t n_strips = (K + strip - 1) / strip;
    const uint32_t P = n_strips > 1 ? (E > 40u ? E : 40u) : E;
    const uint32_t last_cols = K - (n_strips - 1) * strip;
    return (n_strips - 1) * P + E + (last_cols - 1) / (uint32_t)C;
}

// per-step cost: ~96 issue slots of per-step work plus the row update, ~88 per column in the scalar form.  Even C runs the row
// update pairwise on sm_100's packed FP32 instructions (~66 issue slots per column), which pays where the kernel is issue
// bound — the sub-warp classes of short windows (measured: call-methylation windows 2.09 -> 1.97 ms) — and does not in the
// full-warp classes, where the 16 warps of a CTA keep the shared-memory pipe ~80 % busy with the table look-ups' bank conflicts
// (2.6 wavefronts per LDS) and packed C = 10 only adds padding over scalar C = 9 (profiles/r02_k1_variants.md).
#ifndef NPH_EVEN_CELL_COST
#define NPH_EVEN_CELL_COST 66.0f
#endif
NPH_HD float nph_class_cost(uint32_t steps, int C, uint32_t W)
{
    const float cell = ((C & 1) || W == 32u) ? 88.0f : NPH_EVEN_CELL_COST;
    return (float)steps * (96.0f + cell * C) * (W * (1.0f / 32.0f));
}

// returns class index; *steps_out = steps in that class
NPH_HD int nph_choose_class(uint32_t K, uint32_t E, uint32_t* steps_out)
{
    float best = 3.0e38f;
    int best_cls = nph_class_index(NPH_MAX_COLS, 4);
    uint32_t best_steps = 0;
    for (int wi = 0; wi < NPH_NUM_WIDTHS; ++wi) {
        const uint32_t W = nph_class_width(wi);
        for (int C = 1; C <= NPH_MAX_COLS; ++C) {
            const bool fits = K <= W * (uint32_t)C;
            if (nph_class_chained(wi) ? fits : !fits) continue;       // single-strip classes take jobs that fit, the chained class the rest
            const uint32_t steps = nph_class_steps(K, E, C, W);
            const float cost = nph_class_cost(steps, C, W);
            if (cost < best) { best = cost; best_cls = nph_class_index(C, wi); best_steps = steps; }
        }
    }
    *steps_out = best_steps;
    return best_cls;
}

// position inside a class's slice of the schedule: level chunk ascending (so that jobs whose reads land first
// run first while the rest is still crossing PCIe), then steps descending (longest first, lockstep neighbours alike)
NPH_HD uint32_t nph_key_bucket(uint32_t steps, uint32_t chunk)
{
    uint32_t b = steps < 768u ? steps : 768u + (steps - 768u) / 32u;
    if (b >= (uint32_t)NPH_STEP_BUCKETS) b = (uint32_t)NPH_STEP_BUCKETS - 1u;
    if (chunk >= (uint32_t)NPH_CHUNK_BUCKETS) chunk = (uint32_t)NPH_CHUNK_BUCKETS - 1u;
    return chunk * (uint32_t)NPH_STEP_BUCKETS + ((uint32_t)NPH_STEP_BUCKETS - 1u - b);
}
