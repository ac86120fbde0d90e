#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --workload abea > gpurun_out/r02zz_bench_abea.json 2> gpurun_out/r02zz_abea.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:abea_kernel -s 1 -c 1 -o gpurun_out/r02zz_abea \
    python scripts/quick_abea.py 2368 8000 > gpurun_out/r02zz_ncu_abea.log 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/r02zz_bench_abea.json').readline()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'])"
