#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02m_events.txt; : > $O
for mode in fused warm0; do
  case $mode in fused) E="";; warm0) E="NPH_EVENTS_WARMUP=0";; esac
  echo "== pytest $mode" >> $O
  env $E timeout 600 python -m pytest tests/test_gpu_events.py tests/test_gpu_prep.py -q 2>&1 | tail -3 >> $O
done
for w in 64 128; do
  echo "== bench fused warm=$w" >> $O
  NPH_EVENTS_STATS=1 NPH_EVENTS_WARMUP=$w timeout 300 python bench.py --workload events --reads 4096 --steps 5 --warmup 3 2>gpurun_out/r02m_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])" >> $O
  tail -1 gpurun_out/r02m_err.txt >> $O
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ed_ -s 6 -c 2 -o gpurun_out/r02m_events \
    python bench.py --workload events --reads 4096 --steps 1 --warmup 3 > gpurun_out/r02m_ncu.log 2>&1
cat $O
