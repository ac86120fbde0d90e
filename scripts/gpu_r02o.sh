#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02o.txt; : > $O
echo "== pytest events/prep/abea" >> $O
timeout 600 python -m pytest tests/test_gpu_events.py tests/test_gpu_prep.py tests/test_gpu_abea.py -q 2>&1 | tail -3 >> $O
for reads in 4096 512; do for wpr in 1 2 4; do
  echo "== events reads=$reads wpr=$wpr" >> $O
  NPH_EVENTS_STATS=1 NPH_EVENTS_WPR=$wpr timeout 300 python bench.py --workload events --reads $reads --steps 5 --warmup 3 2>gpurun_out/r02o_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])" >> $O
done; done
for v in old new; do
  echo "== abea $v" >> $O
  NPH_LIB_PATH=$PWD/nanopolish_b200/csrc/build/variants/libnph_abea_$v.so timeout 300 python scripts/quick_abea.py 2368 8000 2>&1 | tail -3 >> $O
done
echo "== sanitizer memcheck" >> $O
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py 2>&1 | tail -6 >> $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ed_ -s 6 -c 2 -o gpurun_out/r02o_events \
    python bench.py --workload events --reads 4096 --steps 1 --warmup 3 > gpurun_out/r02o_ncu.log 2>&1
cat $O
