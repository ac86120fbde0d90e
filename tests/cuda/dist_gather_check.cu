// dist_gather_check.cu — multi-device check of libnph_dist.so without Python (test infrastructure): one host thread per visible GPU,
// ncclCommInitAll, every rank holds a different number of records in device memory, nph_dist_gather_records brings them to rank 0;
// then the same with a root whose buffer is too small (every rank must fail alike, nobody may hang), then nph_dist_reduce_sum_f64.
//   nvcc -O2 -std=c++17 -o dist_gather_check dist_gather_check.cu -I../../include -L../../nanopolish_b200 -lnph_dist -lnph -lnccl
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include <nccl.h>
#include "nph_dist.h"

int main()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) { fprintf(stderr, "no CUDA device\n"); return 2; }
    std::vector<ncclComm_t> comms((size_t)n);
    std::vector<int> devs((size_t)n);
    for (int i = 0; i < n; ++i) devs[(size_t)i] = i;
    if (ncclCommInitAll(comms.data(), n, devs.data()) != ncclSuccess) { fprintf(stderr, "ncclCommInitAll failed\n"); return 3; }
    std::vector<int> bad((size_t)n, 0);
    auto worker = [&](int rank) {
        cudaSetDevice(rank);
        nph_ctx* ctx = nullptr;
        if (nph_create(&ctx, rank) != NPH_OK) { bad[(size_t)rank] = 100; return; }
        cudaStream_t st = (cudaStream_t)nph_stream(ctx);
        const size_t mine = (size_t)(1000 + 777 * rank) * 24;                 // records of 24 bytes, a different count per rank
        std::vector<unsigned char> h(mine);
        for (size_t i = 0; i < mine; ++i) h[i] = (unsigned char)((i * 31 + (size_t)rank * 7) & 0xff);
        unsigned char* d_send = nullptr; cudaMalloc((void**)&d_send, mine);
        cudaMemcpyAsync(d_send, h.data(), mine, cudaMemcpyHostToDevice, st);
        size_t total = 0;
        for (int r = 0; r < n; ++r) total += (size_t)(1000 + 777 * r) * 24;
        unsigned char* d_recv = nullptr;
        if (rank == 0) cudaMalloc((void**)&d_recv, total);
        std::vector<uint64_t> counts((size_t)n);
        int rc = nph_dist_gather_records(ctx, comms[(size_t)rank], rank, n, 0, d_send, mine, d_recv, rank == 0 ? total : 0, counts.data());
        if (rc != NPH_OK) bad[(size_t)rank] += 1;
        nph_sync(ctx);
        for (int r = 0; r < n; ++r) if (counts[(size_t)r] != (uint64_t)(1000 + 777 * r) * 24) bad[(size_t)rank] += 1;
        if (rank == 0 && rc == NPH_OK) {
            std::vector<unsigned char> got(total);
            cudaMemcpy(got.data(), d_recv, total, cudaMemcpyDeviceToHost);
            size_t off = 0;
            for (int r = 0; r < n; ++r) {
                const size_t m = (size_t)(1000 + 777 * r) * 24;
                for (size_t i = 0; i < m; ++i) if (got[off + i] != (unsigned char)((i * 31 + (size_t)r * 7) & 0xff)) { bad[0] += 1; break; }
                off += m;
            }
        }
        // a root without enough room: every rank returns NPH_ERR_INVALID, nothing hangs
        rc = nph_dist_gather_records(ctx, comms[(size_t)rank], rank, n, 0, d_send, mine, d_recv, rank == 0 ? total - 1 : 0, nullptr);
        if (rc != NPH_ERR_INVALID) bad[(size_t)rank] += 1;
        nph_sync(ctx);
        // reduce: every rank contributes rank + 1 in 64 doubles
        double* d_v = nullptr; cudaMalloc((void**)&d_v, 64 * sizeof(double));
        double* d_sum = nullptr; cudaMalloc((void**)&d_sum, 64 * sizeof(double));
        std::vector<double> v(64, (double)(rank + 1));
        cudaMemcpyAsync(d_v, v.data(), 64 * sizeof(double), cudaMemcpyHostToDevice, st);
        if (nph_dist_reduce_sum_f64(ctx, comms[(size_t)rank], 0, d_v, d_sum, 64) != NPH_OK) bad[(size_t)rank] += 1;
        nph_sync(ctx);
        if (rank == 0) {
            cudaMemcpy(v.data(), d_sum, 64 * sizeof(double), cudaMemcpyDeviceToHost);
            for (double x : v) if (x != (double)n * (n + 1) / 2) { bad[0] += 1; break; }
        }
        cudaFree(d_send); cudaFree(d_recv); cudaFree(d_v); cudaFree(d_sum);
        nph_destroy(ctx);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < n; ++r) th.emplace_back(worker, r);
    for (auto& t : th) t.join();
    for (int i = 0; i < n; ++i) ncclCommDestroy(comms[(size_t)i]);
    int total_bad = 0;
    for (int b : bad) total_bad += b;
    printf("nph_dist check on %d GPU(s): %s\n", n, total_bad ? "FAILED" : "ok");
    return total_bad ? 1 : 0;
}
