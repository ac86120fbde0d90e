#!/bin/bash
# last check of the tree as the driver will run it: GPU suite, smoke, both bench arms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02w_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02w_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02w_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02w_smoke.log
timeout 500 python bench.py --impl reference > gpurun_out/r02w_bench_reference.json 2> gpurun_out/r02w_ref.err
timeout 600 python bench.py > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_n1.err
tail -3 gpurun_out/r02w_pytest_gpu.log; tail -2 gpurun_out/r02w_smoke.log
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r02w_bench_reference.json').readline()); d=json.loads(open('gpurun_out/r02w_bench_n1.json').readline())
print('ref', r['value'], 'own', d['value'], 'e2e', d['e2e']['value'], 'same config', r['config']==d['config'], 'launches', d['gpu_launches'], d['clocks'])
c=d['configs']['call_methylation']; print('cm', c['value'], c['e2e']['value'], c['e2e']['stage_ms'], c['cpu_baseline'].get('value'), c['cpu_baseline'].get('rows_identical', c['cpu_baseline'].get('tsv_identical')))
PY
