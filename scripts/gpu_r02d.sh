#!/bin/bash
# full GPU suite + bench lines (default line with the call_methylation block, methylation windows) on the current build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest_gpu.log
timeout 300 python bench.py --workload methylation --no-cpu-baseline > gpurun_out/r02d_bench_methylation.json 2> gpurun_out/r02d_bench_methylation.err
timeout 600 python bench.py > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
tail -4 gpurun_out/r02d_pytest_gpu.log; cut -c1-300 gpurun_out/r02d_bench_methylation.json; echo; cat gpurun_out/r02d_bench_n1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('scorereads value %.4g e2e %.4g kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms']))
c=d.get('configs',{}).get('call_methylation'); print(json.dumps(c)[:1500])"; tail -3 gpurun_out/r02d_bench_n1.err
