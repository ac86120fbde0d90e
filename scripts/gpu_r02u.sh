#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02u.txt; : > $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 >> $O
timeout 500 python bench.py --workload variants --no-cpu-baseline > gpurun_out/r02u_bench_variants.json 2> gpurun_out/r02u_v.err
timeout 500 python bench.py --workload call_methylation --no-cpu-baseline > gpurun_out/r02u_bench_call_methylation.json 2> gpurun_out/r02u_cm.err
timeout 500 python bench.py --no-cpu-baseline --no-call-methylation > gpurun_out/r02u_bench_n1.json 2> gpurun_out/r02u_n1.err
python - <<'PY' >> $O
import json
for n in ('r02u_bench_variants','r02u_bench_call_methylation','r02u_bench_n1'):
    try:
        d=json.loads(open('gpurun_out/'+n+'.json').readline())
        print(n, d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('stage_ms'), d['roofline'].get('kernel_ms'))
    except Exception as e: print(n,'ERR',e)
PY
cat $O
