/* np_oracle.h — TEST INFRASTRUCTURE ONLY (see np_oracle.c). */
#ifndef NP_ORACLE_H
#define NP_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/nph.h"   /* POD batch structs only; no product code is linked */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const double* level_mean;
    const double* level_stdv;
    const double* level_log_stdv;
    uint32_t n_states;
} npo_model;

void  npo_init(void);
float npo_logsum(float a, float b);
void  npo_logsum_table(float* out16000);
void  npo_transitions(double events_per_base, double indel_bias, float out10[10]);
void  npo_flank_table(float* out, size_t n);
float npo_log_normal_pdf(float x, float mean, float stdv, float log_stdv);
float npo_drift_scaled_level(const nph_read* read, const float* ev_mean, const double* ev_start_time, uint32_t event_idx);
float npo_log_probability_match(const nph_read* read, const float* ev_mean, const double* ev_start_time,
                                const npo_model* model, uint32_t rank, uint32_t event_idx);

float npo_hmm_score(const nph_read* reads, const float* ev_mean, const double* ev_start_time,
                    const npo_model* models, const uint32_t* kmer_ranks, const nph_hmm_job* job,
                    double indel_bias);
/* Optional: dump the whole forward matrix of one job (rows (E+1) x cols 3*(K+2), row-major). */
float npo_hmm_score_dump(const nph_read* reads, const float* ev_mean, const double* ev_start_time,
                         const npo_model* models, const uint32_t* kmer_ranks, const nph_hmm_job* job,
                         double indel_bias, float* matrix_out);
double npo_hmm_score_batch(const nph_read* reads, const float* ev_mean, const double* ev_start_time,
                           const npo_model* models, const uint32_t* kmer_ranks,
                           const nph_hmm_job* jobs, size_t n_jobs, double indel_bias, int threads,
                           float* scores_out);
float npo_score_set_combine(const float* scores, uint32_t n_alt);

uint32_t npo_hmm_align(const nph_read* reads, const float* ev_mean, const double* ev_start_time,
                       const npo_model* models, const uint32_t* kmer_ranks, const nph_hmm_job* job,
                       double indel_bias, nph_align_state* out, uint32_t cap, int* status);

int64_t npo_abea(const nph_read* reads, const float* ev_mean, const double* ev_start_time,
                 const npo_model* model, const uint32_t* kmer_ranks, const nph_abea_job* job,
                 nph_aligned_pair* pairs_out, nph_abea_result* res);
double npo_abea_batch(const nph_read* reads, const float* ev_mean, const double* ev_start_time,
                      const npo_model* model, const uint32_t* kmer_ranks, const nph_abea_job* jobs,
                      size_t n_jobs, int threads, nph_aligned_pair* pairs_out, nph_abea_result* res);
void npo_mom(const nph_read* reads, const float* ev_mean, const npo_model* model,
             const uint32_t* kmer_ranks, const nph_abea_job* job, double* shift_out, double* scale_out);
long long npo_detect_events(const float* raw, size_t n, const nph_event_params* prm, nph_event* out, size_t cap);
int npo_trim_raw(const float* raw, size_t n, int trim_start, int trim_end, int varseg_chunk, float varseg_thresh,
                 uint32_t* start_out, uint32_t* end_out);
void npo_recalibrate(const nph_read* reads, const float* ev_mean, const npo_model* model, const uint32_t* kmer_ranks,
                     const nph_abea_job* job, const nph_aligned_pair* pairs, uint32_t n_pairs,
                     nph_event_range* b2e, nph_calibration* out);
int npo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
