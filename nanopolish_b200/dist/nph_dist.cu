// nph_dist.cu — libnph_dist.so: the one exchange step of the multi-GPU path behind a C signature (include/nph_dist.h).
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <vector>
#include "../../include/nph_dist.h"

namespace {
struct Scratch {                       // per-thread scratch for the count exchange (a caller drives one device per thread)
    int device = -1;
    uint64_t* d_counts = nullptr;      // 2 * (world + 1) entries: (bytes to send, room when root) per rank
    int cap = 0;
    nph_meth_site* d_recv = nullptr;   // staging of the methylation convenience call
    size_t recv_cap = 0;
};
thread_local Scratch g;

int ensure_counts(int world)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return NPH_ERR_CUDA;
    if (g.device != dev || g.cap < world + 1) {
        if (g.d_counts) cudaFree(g.d_counts);
        g.d_counts = nullptr;
        if (cudaMalloc((void**)&g.d_counts, sizeof(uint64_t) * 2 * (size_t)(world + 1)) != cudaSuccess) return NPH_ERR_NOMEM;
        g.cap = world + 1; g.device = dev;
    }
    return NPH_OK;
}
} // namespace

extern "C" int nph_dist_gather_records(nph_ctx* ctx, void* nccl_comm, int rank, int world, int root,
                                       const void* send_dev, size_t send_bytes,
                                       void* recv_dev, size_t recv_cap, uint64_t* bytes_per_rank_out)
{
    if (!ctx || !nccl_comm || world < 1 || rank < 0 || rank >= world || root < 0 || root >= world) return NPH_ERR_INVALID;
    if (send_bytes && !send_dev) return NPH_ERR_INVALID;
    ncclComm_t comm = (ncclComm_t)nccl_comm;
    cudaStream_t stream = (cudaStream_t)nph_stream(ctx);
    int rc = ensure_counts(world);
    if (rc != NPH_OK) return rc;
    // every rank learns every count AND the root's room, so that all of them take the same decision (post the transfers, or
    // fail together without posting a send nobody receives)
    const uint64_t mine[2] = {(uint64_t)send_bytes, (uint64_t)recv_cap};
    if (cudaMemcpyAsync(g.d_counts + 2 * world, mine, sizeof(mine), cudaMemcpyHostToDevice, stream) != cudaSuccess) return NPH_ERR_CUDA;
    if (ncclAllGather(g.d_counts + 2 * world, g.d_counts, 2, ncclUint64, comm, stream) != ncclSuccess) return NPH_ERR_CUDA;
    std::vector<uint64_t> both(2 * (size_t)world), counts((size_t)world);
    if (cudaMemcpyAsync(both.data(), g.d_counts, sizeof(uint64_t) * 2 * (size_t)world, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return NPH_ERR_CUDA;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return NPH_ERR_CUDA;
    for (int r = 0; r < world; ++r) counts[(size_t)r] = both[2 * (size_t)r];
    if (bytes_per_rank_out) for (int r = 0; r < world; ++r) bytes_per_rank_out[r] = counts[(size_t)r];
    uint64_t total = 0;
    for (uint64_t c : counts) total += c;
    const bool fits = total <= both[2 * (size_t)root + 1];
    if (!fits) return NPH_ERR_INVALID;                     // on every rank alike
    if (rank == root && total && !recv_dev) return NPH_ERR_INVALID;
    if (ncclGroupStart() != ncclSuccess) return NPH_ERR_CUDA;
    if (rank == root) {
        uint64_t off = 0;
        for (int r = 0; r < world; ++r) {
            if (counts[(size_t)r]) {
                if (r == root) {
                    if (cudaMemcpyAsync((char*)recv_dev + off, send_dev, (size_t)counts[(size_t)r], cudaMemcpyDeviceToDevice, stream) != cudaSuccess) { ncclGroupEnd(); return NPH_ERR_CUDA; }
                } else if (ncclRecv((char*)recv_dev + off, (size_t)counts[(size_t)r], ncclUint8, r, comm, stream) != ncclSuccess) { ncclGroupEnd(); return NPH_ERR_CUDA; }
            }
            off += counts[(size_t)r];
        }
    } else if (rank != root && send_bytes) {
        if (ncclSend(send_dev, send_bytes, ncclUint8, root, comm, stream) != ncclSuccess) { ncclGroupEnd(); return NPH_ERR_CUDA; }
    }
    if (ncclGroupEnd() != ncclSuccess) return NPH_ERR_CUDA;
    return NPH_OK;
}

extern "C" int nph_dist_gather_methylation_sites(nph_ctx* ctx, void* nccl_comm, int rank, int world, int root,
                                                 nph_meth_site* sites_out, size_t sites_cap, uint64_t* n_sites_per_rank_out)
{
    if (!ctx) return NPH_ERR_INVALID;
    const nph_meth_site* d_sites = nullptr;
    uint64_t n_sites = 0;
    int rc = nph_methylation_sites_dev(ctx, &d_sites, &n_sites);
    if (rc != NPH_OK) return rc;
    if (rank == root) {
        if (sites_cap && !sites_out) return NPH_ERR_INVALID;
        if (g.recv_cap < sites_cap) {
            if (g.d_recv) cudaFree(g.d_recv);
            g.d_recv = nullptr; g.recv_cap = 0;
            if (sites_cap && cudaMalloc((void**)&g.d_recv, sizeof(nph_meth_site) * sites_cap) != cudaSuccess) return NPH_ERR_NOMEM;
            g.recv_cap = sites_cap;
        }
    }
    std::vector<uint64_t> bytes((size_t)world);
    rc = nph_dist_gather_records(ctx, nccl_comm, rank, world, root, d_sites, (size_t)n_sites * sizeof(nph_meth_site),
                                 rank == root ? g.d_recv : nullptr, rank == root ? sites_cap * sizeof(nph_meth_site) : 0, bytes.data());
    if (n_sites_per_rank_out) for (int r = 0; r < world; ++r) n_sites_per_rank_out[r] = bytes[(size_t)r] / sizeof(nph_meth_site);
    if (rc != NPH_OK) return rc;
    cudaStream_t stream = (cudaStream_t)nph_stream(ctx);
    if (rank == root) {
        uint64_t total = 0;
        for (uint64_t b : bytes) total += b;
        if (total && cudaMemcpyAsync(sites_out, g.d_recv, (size_t)total, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return NPH_ERR_CUDA;
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess) return NPH_ERR_CUDA;
    return NPH_OK;
}

extern "C" int nph_dist_reduce_sum_f64(nph_ctx* ctx, void* nccl_comm, int root, const double* send_dev, double* recv_dev, size_t count)
{
    if (!ctx || !nccl_comm || (count && !send_dev)) return NPH_ERR_INVALID;
    if (count == 0) return NPH_OK;
    if (ncclReduce(send_dev, recv_dev, count, ncclDouble, ncclSum, root, (ncclComm_t)nccl_comm, (cudaStream_t)nph_stream(ctx)) != ncclSuccess) return NPH_ERR_CUDA;
    return NPH_OK;
}
