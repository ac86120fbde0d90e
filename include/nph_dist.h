/* nph_dist.h — multi-GPU result gather for C / C++ callers of libnph.so (libnph_dist.so).
 *
 * reads are independent for every kernel of the path, so a caller shards a BamProcessor batch over the
 * GPUs of a box (one nph_ctx per device, one host thread or process per device) and there is no data-path collective during
 * compute; what remains is ONE exchange per batch — every rank's result records (nph_meth_site rows, per-job scores, eventalign
 * records: fixed-size PODs, a different count on every rank) travel to the root over NVLink.  The reference has no counterpart
 * (it is one process with OpenMP threads, src/common/nanopolish_bam_processor.cpp:99); this is the "single NCCL gather" of
 * BASELINE.json's north_star, behind a plain C signature so that the C++ host shards without Python.
 *
 * libnph_dist.so links NCCL (ncclComm_t crosses the boundary as void*); libnph.so itself stays free of it.  The Python
 * driver (bench.py, nanopolish_b200/dist.py) uses torch.distributed for the same exchange.
 */
#ifndef NPH_DIST_H
#define NPH_DIST_H

#include <stddef.h>
#include <stdint.h>
#include "nph.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Variable-length gather of device-resident records to `root`, queued on the context's stream behind whatever produced them:
 *   1. a world-sized all-gather of the byte counts (8 bytes per rank),
 *   2. one grouped send/recv: rank r's send_bytes bytes land at recv_dev + sum of the counts of the ranks before it.
 * send_dev: device pointer on ctx's device (may be NULL when send_bytes == 0), e.g. from nph_methylation_sites_dev.
 * recv_dev / recv_cap: root only — device buffer and its size; if the total does not fit, EVERY rank returns NPH_ERR_INVALID and
 * nothing is sent (the root's room travels with the counts).
 * bytes_per_rank_out: optional, `world` entries, filled on EVERY rank (the counts are all-gathered).
 * The call synchronises the stream once (the counts are needed on the host to post the receives). */
int nph_dist_gather_records(nph_ctx* ctx, void* nccl_comm, int rank, int world, int root,
                            const void* send_dev, size_t send_bytes,
                            void* recv_dev, size_t recv_cap, uint64_t* bytes_per_rank_out);

/* Convenience over the above for call-methylation: gathers the site records of the most recent nph_methylation_run of every
 * rank to the root's HOST buffer sites_out (root only; sites_cap records), in rank order; records' `record` field stays
 * rank-local (the caller knows which reads it gave to which rank).  n_sites_per_rank_out: `world` entries, every rank. */
int nph_dist_gather_methylation_sites(nph_ctx* ctx, void* nccl_comm, int rank, int world, int root,
                                      nph_meth_site* sites_out, size_t sites_cap, uint64_t* n_sites_per_rank_out);

/* ncclReduce(sum) of per-candidate score sums to the root (config 5: every rank scores its reads against all candidates,
 * SURVEY.md 8e), double precision, device buffers, on the context's stream. */
int nph_dist_reduce_sum_f64(nph_ctx* ctx, void* nccl_comm, int root, const double* send_dev, double* recv_dev, size_t count);

#ifdef __cplusplus
}
#endif
#endif /* NPH_DIST_H */
