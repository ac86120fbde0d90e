#!/bin/bash
# GPU parity suite + eventalign chain timing + sanitizer pass over the chain kernel.  bash scripts/gpu_round2.sh [tag]
tag=${1:-r01c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
timeout 300 python scripts/quick_eventalign.py 256 4000 32 > gpurun_out/${tag}_eventalign_256.json 2> gpurun_out/${tag}_eventalign_256.err
timeout 500 python scripts/quick_eventalign.py 2368 4000 32 > gpurun_out/${tag}_eventalign_2368.json 2> gpurun_out/${tag}_eventalign_2368.err
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_eventalign.py -m gpu -x -q -k "on_device or abi or falls_back" > gpurun_out/${tag}_sanitizer_chain.log 2>&1
tail -5 gpurun_out/${tag}_pytest_gpu.log; cat gpurun_out/${tag}_eventalign_256.json gpurun_out/${tag}_eventalign_2368.json; tail -4 gpurun_out/${tag}_eventalign_256.err; tail -6 gpurun_out/${tag}_sanitizer_chain.log
