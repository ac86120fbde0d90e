#!/bin/bash
# End-of-round evidence: GPU parity suite, smoke, bench lines (ours + reference arm, every workload), launch list, ncu captures of
# the kernels that changed this round, host-path timings, sanitizer.  Usage: bash scripts/gpu_final.sh [tag]
tag=${1:-r02z}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $O/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${tag}_smoke.log
timeout 500 python bench.py --impl reference > $O/${tag}_bench_reference.json 2> $O/${tag}_bench_reference.err
timeout 600 python bench.py > $O/${tag}_bench_n1.json 2> $O/${tag}_bench_n1.err
timeout 500 python bench.py --workload call_methylation > $O/${tag}_bench_call_methylation.json 2> $O/${tag}_bench_call_methylation.err
timeout 500 python bench.py --workload variants > $O/${tag}_bench_variants.json 2> $O/${tag}_bench_variants.err
timeout 300 python bench.py --workload methylation --no-cpu-baseline > $O/${tag}_bench_methylation.json 2> $O/${tag}_bench_methylation.err
timeout 400 python bench.py --workload abea > $O/${tag}_bench_abea.json 2> $O/${tag}_bench_abea.err
timeout 300 python bench.py --workload events --reads 4096 > $O/${tag}_bench_events.json 2> $O/${tag}_bench_events.err
timeout 300 python bench.py --workload prologue > $O/${tag}_bench_prologue.json 2> $O/${tag}_bench_prologue.err
timeout 400 python bench.py --workload eventalign --reads 2368 > $O/${tag}_bench_eventalign.json 2> $O/${tag}_bench_eventalign.err
timeout 400 python scripts/quick_methylation.py 4096 4000 > $O/${tag}_methylation_host.json 2> $O/${tag}_methylation_host.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${tag}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ed_ -s 6 -c 2 -o $O/${tag}_events \
    python bench.py --workload events --reads 4096 --steps 1 --warmup 3 > $O/${tag}_ncu_events.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:abea_kernel -s 1 -c 1 -o $O/${tag}_abea \
    python scripts/quick_abea.py 2368 8000 > $O/${tag}_ncu_abea.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py > $O/${tag}_sanitizer.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py >> $O/${tag}_sanitizer.txt 2>&1
tail -3 $O/${tag}_pytest_gpu.log; tail -2 $O/${tag}_smoke.log
python - <<PY
import json
def load(n):
    try: return json.loads(open('$O/${tag}_'+n+'.json').readline())
    except Exception as e: return {'error': repr(e)}
for n in ('bench_reference','bench_n1','bench_call_methylation','bench_variants','bench_methylation','bench_abea','bench_events','bench_prologue','bench_eventalign'):
    d=load(n)
    print(n, d.get('value'), d.get('unit'), 'e2e', (d.get('e2e') or {}).get('value'), 'roofline', (d.get('roofline') or {}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('error',''))
d=load('bench_n1'); c=(d.get('configs') or {}).get('call_methylation') or {}
print('configs.call_methylation', c.get('value'), (c.get('e2e') or {}).get('value'), (c.get('cpu_baseline') or {}).get('value'))
PY
cut -c1-700 $O/${tag}_methylation_host.json; grep -c "ERROR SUMMARY: 0 errors" $O/${tag}_sanitizer.txt; grep "ERROR SUMMARY" $O/${tag}_sanitizer.txt
