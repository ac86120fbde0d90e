#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_methylation.py tests/test_host_methylation.py -q > gpurun_out/r02g_pytest.log 2>&1; tail -3 gpurun_out/r02g_pytest.log
timeout 900 python bench.py --workload variants --region 100000 --steps 3 > gpurun_out/r02g_bench_variants.json 2> gpurun_out/r02g_bench_variants.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02g_bench_variants.json'))
    print('variants value %.4g ms/step %.2f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))
    print({k:d['config'][k] for k in ('rounds','reference_dp_rows_per_step','our_dp_rows_per_step','jobs_per_step','jobs_without_early_exit','mean_event_sequences_per_position')})
    print(d['roofline']['kernel_ms'], d.get('cpu_baseline'))
except Exception as e:
    print('bench variants failed', e)
PY
tail -5 gpurun_out/r02g_bench_variants.err
timeout 600 python bench.py --workload call_methylation > gpurun_out/r02g_bench_call_methylation.json 2> gpurun_out/r02g_bench_call_methylation.err
python - <<'PY'
import json
c=json.load(open('gpurun_out/r02g_bench_call_methylation.json'))
print('call_methylation value %.4g e2e %.4g' % (c['value'], c['e2e']['value']), c['e2e']['stage_ms'], c['e2e']['ms_per_step'])
PY
timeout 300 python scripts/quick_methylation.py 4096 4000 2>/dev/null | cut -c1-700
