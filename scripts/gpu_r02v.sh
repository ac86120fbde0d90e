#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02v.txt; : > $O
for v in default ed_r2c4 ed_r2c5; do
  if [ $v = default ]; then L=""; else L="NPH_LIB_PATH=$PWD/nanopolish_b200/csrc/build/variants/libnph_$v.so"; fi
  echo "== pytest $v" >> $O
  env $L timeout 600 python -m pytest tests/test_gpu_events.py -q 2>&1 | tail -2 >> $O
  for reads in 4096 8192 512; do
    echo "== events $v reads=$reads" >> $O
    env $L timeout 300 python bench.py --workload events --reads $reads --steps 5 --warmup 3 2>gpurun_out/r02v_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'])" >> $O
  done
done
cat $O
timeout 400 python scripts/quick_methylation.py 4096 4000 > gpurun_out/r02v_methylation_host.json 2> gpurun_out/r02v_methylation_host.err
cut -c1-600 gpurun_out/r02v_methylation_host.json; tail -2 gpurun_out/r02v_methylation_host.err
