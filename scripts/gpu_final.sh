#!/bin/bash
# End-of-round evidence: GPU parity suite, smoke, bench lines (ours + reference arm), launch list, host-path timings.
tag=${1:-r01z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 400 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 300 python scripts/quick_methylation.py 512 4000 > gpurun_out/${tag}_meth_512.json 2> gpurun_out/${tag}_meth_512.err
timeout 400 python scripts/quick_methylation.py 4096 4000 > gpurun_out/${tag}_meth_4096.json 2> gpurun_out/${tag}_meth_4096.err
timeout 300 python bench.py --workload methylation --no-cpu-baseline > gpurun_out/${tag}_bench_methylation.json 2> gpurun_out/${tag}_bench_methylation.err
tail -3 gpurun_out/${tag}_pytest_gpu.log; tail -1 gpurun_out/${tag}_smoke.log; cut -c1-300 gpurun_out/${tag}_bench_n1.json; cat gpurun_out/${tag}_meth_512.json gpurun_out/${tag}_meth_4096.json; cut -c1-250 gpurun_out/${tag}_bench_methylation.json
