// ubench_f32x2.cu — issue-rate probe for sm_100a packed FP32 (FADD2/FMUL2/FFMA2) against scalar FADD/FMUL/FFMA,
// alone and mixed with ALU-pipe work (FMNMX), the mix K1's log-sum uses.  Development aid.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_f32x2 ubench_f32x2.cu && ./ubench_f32x2
#include <cstdio>
#include <cuda_runtime.h>

#define PK(a, lo, hi) asm volatile("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(lo), "f"(hi))
#define UP(lo, hi, a) asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a))

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float* out, int iters, float seed)
{
    // 8 independent chains per thread so latency never binds; each chain element = one FP32 lane-op
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; b[i] = 1.0f + 1e-7f * i; }
    unsigned long long p[4], q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { PK(p[i], a[2 * i], a[2 * i + 1]); PK(q[i], b[2 * i], b[2 * i + 1]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            if (MODE == 0) {          // 8 scalar FADD
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            } else if (MODE == 1) {   // 4 FADD2 (same lane-ops)
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
            } else if (MODE == 2) {   // 8 scalar FFMA
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b[i]));
            } else if (MODE == 3) {   // 4 FFMA2
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(q[i]));
            } else if (MODE == 4) {   // 8 FADD + 8 FMNMX (alu pipe)
#pragma unroll
                for (int i = 0; i < 8; ++i) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i])); asm volatile("max.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(a[i])); }
            } else if (MODE == 5) {   // 4 FADD2 + 8 FMNMX on the unpacked halves (forces pack/unpack-free use: halves are registers)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
                    float lo, hi, ql, qh; UP(lo, hi, p[i]); UP(ql, qh, q[i]);
                    asm volatile("max.f32 %0, %0, %1;" : "+f"(ql) : "f"(lo));
                    asm volatile("max.f32 %0, %0, %1;" : "+f"(qh) : "f"(hi));
                    PK(q[i], ql, qh);
                }
            } else if (MODE == 6) {   // 8 scalar FMUL
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            } else if (MODE == 7) {   // 4 FMUL2
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
            } else if (MODE == 8) {   // 8 FADD.RM scalar
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            } else if (MODE == 9) {   // 4 FADD2.RM
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) { float lo, hi; UP(lo, hi, p[i]); s += lo + hi; UP(lo, hi, q[i]); s += lo + hi; }
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, int lane_ops_per_rep)
{
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, grid = 148;
    probe<MODE><<<grid, 512>>>(d, 100, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<grid, 512>>>(d, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)grid * 512 * iters * 8.0 * lane_ops_per_rep;
    // per SMSP per cycle at 1.965 GHz: lane-ops / (148*4) / (ms*1e-3*1.965e9)
    printf("%-28s %8.3f ms  %7.2f FP32 lane-ops/clk/SMSP (at 1965 MHz)  err=%s\n", name, ms,
           lane_ops / (148.0 * 4) / (ms * 1e-3 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main()
{
    run<0>("8 FADD", 8); run<1>("4 FADD2", 8); run<2>("8 FFMA", 8); run<3>("4 FFMA2", 8);
    run<6>("8 FMUL", 8); run<7>("4 FMUL2", 8); run<8>("8 FADD.RM", 8); run<9>("4 FADD2.RM", 8);
    run<4>("8 FADD + 8 FMNMX", 8); run<5>("4 FADD2 + 8 FMNMX", 8);
    return 0;
}
