#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest_gpu.log
timeout 600 python bench.py --workload call_methylation > gpurun_out/r02e_bench_call_methylation.json 2> gpurun_out/r02e_bench_call_methylation.err
timeout 300 python scripts/quick_methylation.py 4096 4000 > gpurun_out/r02e_meth_4096.json 2> gpurun_out/r02e_meth_4096.err
tail -4 gpurun_out/r02e_pytest_gpu.log; cat gpurun_out/r02e_bench_call_methylation.json | python -c "
import sys,json
c=json.loads(sys.stdin.read()); print('value %.4g ms/step %.3f' % (c['value'], c['ms_per_step'])); print(json.dumps(c['e2e'])); print(json.dumps(c.get('cpu_baseline'))[:600]); print(json.dumps(c['roofline'])[:400])"; tail -2 gpurun_out/r02e_bench_call_methylation.err; cat gpurun_out/r02e_meth_4096.json | cut -c1-700; tail -5 gpurun_out/r02e_meth_4096.err
