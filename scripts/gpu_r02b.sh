#!/bin/bash
# round-2 second GPU call: full GPU suite on the packed-FP32 K1, bench lines, ncu of the packed kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
timeout 300 python bench.py --workload methylation --no-cpu-baseline > gpurun_out/r02b_bench_methylation.json 2> gpurun_out/r02b_bench_methylation.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hmm_forward -s 2 -c 2 -o gpurun_out/r02b_fwd_packed \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02b_ncu_full.log 2>&1
tail -5 gpurun_out/r02b_pytest_gpu.log; cut -c1-400 gpurun_out/r02b_bench_n1.json; echo; cut -c1-300 gpurun_out/r02b_bench_methylation.json
