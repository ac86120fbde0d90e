#!/bin/bash
# A/B in one session: scorereads with the base-code prologue (codes), the same build driven through the rank form, and a build with the code path compiled out
mkdir -p gpurun_out; out=gpurun_out/r02i_codes_ab.txt; : > $out
run() { # label, lib, env
  for i in 1 2; do
  env $3 NPH_LIB_PATH=$PWD/nanopolish_b200/csrc/build/variants/libnph_$2.so timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-call-methylation 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'value=%.4g' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'], 'e2e=%.4g' % d['e2e']['value'])" >> $out 2>&1
  done
}
run cur_codes cur "X=1"
run cur_ranks cur "NPH_BENCH_RANKS=1"
run nocodes_ranks nocodes "NPH_BENCH_RANKS=1"
cat $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 1 --workload call_methylation --steps 5 > /dev/null 2>&1
