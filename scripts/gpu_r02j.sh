#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r02j_bench_ref.json 2> gpurun_out/r02j_bench_ref.err
tail -4 gpurun_out/r02j_pytest_gpu.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench_n1.json'))
print('scorereads value %.4g e2e %.4g h2d %d kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['roofline']['kernel_ms']))
c=d['configs']['call_methylation']; print('call_methylation value %.4g e2e %.4g' % (c['value'], c['e2e']['value']), c['e2e']['stage_ms'], c.get('cpu_baseline',{}).get('value'))
r=json.load(open('gpurun_out/r02j_bench_ref.json')); print('ref', r['value'], r['config'])
print('own config', d['config'])
PY
