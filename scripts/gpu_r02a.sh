#!/bin/bash
# round-2 first GPU call: packed-FP32 issue probe, ncu of the call-methylation window classes, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
timeout 120 nanopolish_b200/csrc/build/ubench_f32x2 > gpurun_out/r02a_ubench_f32x2.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02a_launches_methylation.csv \
    python bench.py --workload methylation --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_ncu_meth_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hmm_forward -s 10 -c 10 -o gpurun_out/r02a_meth_fwd \
    python bench.py --workload methylation --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_ncu_meth_full.log 2>&1
timeout 300 python bench.py --workload methylation --no-cpu-baseline > gpurun_out/r02a_bench_methylation.json 2> gpurun_out/r02a_bench_methylation.err
cat gpurun_out/r02a_ubench_f32x2.txt; cut -c1-300 gpurun_out/r02a_bench_methylation.json; ls -la gpurun_out | tail -8
