#!/bin/bash
# N=2 check of every multi-GPU line (default bench with the call-methylation block, reference arm under torchrun, variants)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $T bench.py --gpus 2 > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err; echo "rc=$?"
timeout 400 $T bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r02q_bench_ref_n2.json 2> gpurun_out/r02q_bench_ref_n2.err; echo "rc=$?"
timeout 500 $T bench.py --gpus 2 --workload variants --no-cpu-baseline > gpurun_out/r02q_bench_variants_n2.json 2> gpurun_out/r02q_bench_variants_n2.err; echo "rc=$?"
timeout 200 ./tests/cuda/dist_gather_check; echo "dist rc=$?"
python - <<'PY'
import json
for n in ('bench_n2','bench_ref_n2','bench_variants_n2'):
    try:
        d=json.loads(open('gpurun_out/r02q_'+n+'.json').readline())
        print(n, d.get('n_gpus'), d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'ms', d.get('ms_per_step'))
        c=(d.get('configs') or {}).get('call_methylation')
        if c: print('   call_methylation', c.get('value'), (c.get('e2e') or {}).get('value'), c.get('ms_per_step'))
    except Exception as e: print(n, 'ERR', e)
PY
tail -3 gpurun_out/r02q_bench_n2.err
