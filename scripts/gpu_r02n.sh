#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02n_events.txt; : > $O
for mode in "X=1" "NPH_EVENTS_WPR=1" "NPH_EVENTS_WPR=2" "NPH_EVENTS_WPR=4" "NPH_EVENTS_WARMUP=0 NPH_EVENTS_WPR=2" "NPH_EVENTS_WARMUP=0 NPH_EVENTS_WPR=4"; do
  echo "== pytest $mode" >> $O
  env $mode timeout 600 python -m pytest tests/test_gpu_events.py tests/test_gpu_prep.py -q 2>&1 | tail -3 >> $O
done
for reads in 4096 512; do for wpr in 1 2 4; do
  echo "== bench reads=$reads wpr=$wpr" >> $O
  NPH_EVENTS_STATS=1 NPH_EVENTS_WPR=$wpr timeout 300 python bench.py --workload events --reads $reads --steps 5 --warmup 3 2>gpurun_out/r02n_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])" >> $O
  tail -1 gpurun_out/r02n_err.txt >> $O
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ed_ -s 6 -c 2 -o gpurun_out/r02n_events \
    python bench.py --workload events --reads 4096 --steps 1 --warmup 3 > gpurun_out/r02n_ncu.log 2>&1
cat $O
