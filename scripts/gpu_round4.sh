#!/bin/bash
# host-thread scaling of the eventalign host path, bench eventalign with pinned e2e, sanitizer smoke (memcheck)
tag=${1:-r01e}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/${tag}_host.txt 2>&1; lscpu | grep -E "Model name|Thread|Core|Socket|^CPU\(s\)" >> gpurun_out/${tag}_host.txt
for t in 4 16 32 64; do
  NPH_HOST_THREADS=$t timeout 300 python scripts/quick_eventalign.py 1024 4000 8 > gpurun_out/${tag}_ea_t${t}.json 2> gpurun_out/${tag}_ea_t${t}.err
done
timeout 400 python bench.py --workload eventalign --reads 2368 > gpurun_out/${tag}_bench_eventalign.json 2> gpurun_out/${tag}_bench_eventalign.err
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py > gpurun_out/${tag}_sanitizer.log 2>&1
cat gpurun_out/${tag}_host.txt; for t in 4 16 32 64; do cat gpurun_out/${tag}_ea_t${t}.json; echo; done; cat gpurun_out/${tag}_bench_eventalign.json; tail -3 gpurun_out/${tag}_bench_eventalign.err; tail -4 gpurun_out/${tag}_sanitizer.log
