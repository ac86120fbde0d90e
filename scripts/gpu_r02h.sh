#!/bin/bash
# 2-GPU validation: the C++ NCCL gather checker, the default bench line (scorereads + call_methylation block) and the variants workload
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 120 tests/cuda/dist_gather_check > gpurun_out/r02h_dist_check.txt 2>&1; cat gpurun_out/r02h_dist_check.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02h_bench_n2.json'))
    print('n2 scorereads value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
    c=d['configs']['call_methylation']; print('n2 call_methylation', {k:c.get(k) for k in ('value','ms_per_step','error')}, c.get('e2e',{}).get('value'))
except Exception as e: print('n2 parse failed', e)
PY
tail -3 gpurun_out/r02h_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload variants --region 50000 --steps 2 --warmup 1 \
    > gpurun_out/r02h_bench_variants_n2.json 2> gpurun_out/r02h_bench_variants_n2.err; echo "variants n2 rc=$?"; cut -c1-300 gpurun_out/r02h_bench_variants_n2.json; tail -3 gpurun_out/r02h_bench_variants_n2.err
