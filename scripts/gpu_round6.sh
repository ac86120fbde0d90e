#!/bin/bash
tag=${1:-r01g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py --workload eventalign --reads 2368 --no-cpu-baseline > gpurun_out/${tag}_bench_eventalign.json 2> gpurun_out/${tag}_bench_eventalign.err
timeout 200 python scripts/quick_viterbi.py 500 > gpurun_out/${tag}_viterbi.log 2>&1
tail -6 gpurun_out/${tag}_pytest_gpu.log; cut -c1-400 gpurun_out/${tag}_bench_eventalign.json; tail -2 gpurun_out/${tag}_viterbi.log
