#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_events.py tests/test_gpu_prep.py -q 2>&1 | tail -3
timeout 300 python bench.py --workload events --reads 4096 > gpurun_out/r02x_bench_events.json 2> gpurun_out/r02x_ev.err
python -c "
import json
d=json.loads(open('gpurun_out/r02x_bench_events.json').readline()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e'], d['roofline']['conversion_unit'])"
