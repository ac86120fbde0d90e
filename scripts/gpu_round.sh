#!/bin/bash
# One gpurun call's worth of evidence: GPU parity suite, eventalign timing, bench lines (ours + reference arm), launch list.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
timeout 400 python scripts/quick_eventalign.py 256 4000 32 > gpurun_out/${tag}_eventalign.json 2> gpurun_out/${tag}_eventalign.err
timeout 400 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 400 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
tail -3 gpurun_out/${tag}_pytest_gpu.log; cat gpurun_out/${tag}_eventalign.json; cat gpurun_out/${tag}_bench_n1.json
