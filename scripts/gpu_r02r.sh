#!/bin/bash
# N=8 check of the default bench line (the driver's scaling run) — scorereads + the call-methylation block (100 000 reads in all)
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $T bench.py --gpus $N > gpurun_out/r02r_bench_n$N.json 2> gpurun_out/r02r_bench_n$N.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02r_bench_n$N.json').readline())
    print(d.get('n_gpus'), d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'ms', d.get('ms_per_step'), d.get('clocks'))
    c=(d.get('configs') or {}).get('call_methylation')
    if c: print('   call_methylation', c.get('value'), (c.get('e2e') or {}), c.get('ms_per_step'), c.get('config',{}).get('reads_total'))
except Exception as e: print('ERR', e)
PY
tail -5 gpurun_out/r02r_bench_n$N.err
