#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02s.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_methylation.py tests/test_host_methylation.py tests/test_abi.py -q -m gpu 2>&1 | tail -8 >> $O
timeout 500 python bench.py --workload call_methylation > gpurun_out/r02s_bench_call_methylation.json 2> gpurun_out/r02s_cm.err
NPH_METH_HOST_TSV=1 timeout 500 python bench.py --workload call_methylation --no-cpu-baseline > gpurun_out/r02s_bench_call_methylation_hosttsv.json 2> gpurun_out/r02s_cm2.err
python - <<'PY' >> $O
import json
for n in ('r02s_bench_call_methylation','r02s_bench_call_methylation_hosttsv'):
    try:
        d=json.loads(open('gpurun_out/'+n+'.json').readline())
        print(n, d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('stage_ms'), d['e2e'].get('ms_per_step'), d['e2e'].get('d2h_bytes_per_step'))
    except Exception as e: print(n,'ERR',e)
PY
tail -3 gpurun_out/r02s_cm.err >> $O
cat $O
