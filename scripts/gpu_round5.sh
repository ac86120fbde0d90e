#!/bin/bash
tag=${1:-r01f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
timeout 400 python bench.py --workload eventalign --reads 2368 > gpurun_out/${tag}_bench_eventalign.json 2> gpurun_out/${tag}_bench_eventalign.err
timeout 300 python scripts/quick_eventalign.py 1024 4000 8 > gpurun_out/${tag}_eventalign_1024.json 2> gpurun_out/${tag}_eventalign_1024.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:eventalign_chain -s 1 -c 1 -f -o gpurun_out/${tag}_chain \
    python bench.py --workload eventalign --reads 2368 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_chain.log 2>&1
timeout 200 python scripts/quick_viterbi.py 500 > gpurun_out/${tag}_viterbi.log 2>&1
tail -4 gpurun_out/${tag}_pytest_gpu.log; cat gpurun_out/${tag}_bench_eventalign.json gpurun_out/${tag}_eventalign_1024.json; tail -2 gpurun_out/${tag}_viterbi.log
