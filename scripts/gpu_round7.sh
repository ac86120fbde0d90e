#!/bin/bash
tag=${1:-r01h}
mkdir -p gpurun_out
timeout 300 python scripts/quick_methylation.py 512 4000 > gpurun_out/${tag}_meth_512.json 2> gpurun_out/${tag}_meth_512.err
timeout 500 python scripts/quick_methylation.py 4096 4000 > gpurun_out/${tag}_meth_4096.json 2> gpurun_out/${tag}_meth_4096.err
NPH_HOST_THREADS=1 timeout 300 python scripts/quick_methylation.py 512 4000 > gpurun_out/${tag}_meth_512_t1.json 2> gpurun_out/${tag}_meth_512_t1.err
cat gpurun_out/${tag}_meth_512.json gpurun_out/${tag}_meth_4096.json gpurun_out/${tag}_meth_512_t1.json; tail -3 gpurun_out/${tag}_meth_4096.err
