#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
timeout 300 python scripts/quick_methylation.py 4096 4000 > gpurun_out/r02f_meth_4096.json 2> gpurun_out/r02f_meth_4096.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hmm_forward -s 10 -c 10 -o gpurun_out/r02f_meth_fwd \
    python bench.py --workload methylation --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02f_ncu_meth_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hmm_forward -s 2 -c 2 -o gpurun_out/r02f_score_fwd \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-call-methylation > gpurun_out/r02f_ncu_score_full.log 2>&1
tail -4 gpurun_out/r02f_pytest_gpu.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench_n1.json'))
print('scorereads value %.4g e2e %.4g h2d %d kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['roofline']['kernel_ms']))
c=d['configs']['call_methylation']; print('call_methylation value %.4g e2e %.4g' % (c['value'], c['e2e']['value']), c['e2e']['stage_ms'], c.get('cpu_baseline',{}).get('value'))
PY
cat gpurun_out/r02f_meth_4096.json | cut -c1-900; tail -3 gpurun_out/r02f_meth_4096.err
