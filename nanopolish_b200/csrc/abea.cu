// abea.cu — K2: adaptive banded event alignment (placeholder until the kernel lands)
#include "nph_internal.cuh"

int nph_launch_abea(nph_ctx*) { return NPH_ERR_UNSUPPORTED; }

extern "C" {
int nph_abea_batch(nph_ctx*, const nph_read*, size_t, const float*, const double*, size_t, const uint32_t*, size_t,
                   const nph_abea_job*, size_t, uint32_t, nph_aligned_pair*, size_t, nph_abea_result*) { return NPH_ERR_UNSUPPORTED; }
int nph_abea_jobs_load(nph_ctx*, const uint32_t*, size_t, const nph_abea_job*, size_t, uint32_t, size_t) { return NPH_ERR_UNSUPPORTED; }
int nph_abea_run(nph_ctx*) { return NPH_ERR_UNSUPPORTED; }
int nph_abea_fetch(nph_ctx*, nph_aligned_pair*, size_t, nph_abea_result*, size_t) { return NPH_ERR_UNSUPPORTED; }
int nph_mom_batch(nph_ctx*, const nph_read*, size_t, const float*, size_t, const uint32_t*, size_t,
                  const nph_abea_job*, size_t, uint32_t, double*) { return NPH_ERR_UNSUPPORTED; }
int nph_hmm_align_batch(nph_ctx*, const nph_read*, size_t, const float*, const double*, size_t, const uint32_t*, size_t,
                        const nph_hmm_job*, size_t, double, nph_align_state*, const uint64_t*, uint32_t*, float*) { return NPH_ERR_UNSUPPORTED; }
}
