#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02p.txt; : > $O
echo "== pytest events/prep" >> $O
timeout 600 python -m pytest tests/test_gpu_events.py tests/test_gpu_prep.py -q 2>&1 | tail -3 >> $O
for v in ed5 ed6; do for reads in 4096 512; do
  echo "== events $v reads=$reads" >> $O
  NPH_LIB_PATH=$PWD/nanopolish_b200/csrc/build/variants/libnph_$v.so timeout 300 python bench.py --workload events --reads $reads --steps 5 --warmup 3 2>gpurun_out/r02p_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])" >> $O
done; done
cat $O
