#!/bin/bash
mkdir -p gpurun_out
./scripts/ubench_cvt > gpurun_out/r02l_ubench_cvt.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ed_ -s 6 -c 2 -o gpurun_out/r02l_events \
    python bench.py --workload events --reads 4096 --steps 1 --warmup 3 > gpurun_out/r02l_ncu.log 2>&1
cat gpurun_out/r02l_ubench_cvt.txt; tail -3 gpurun_out/r02l_ncu.log | cut -c1-300
