#!/bin/bash
tag=${1:-r01i}
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:abea_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_abea python scripts/quick_abea.py 2368 8000 > gpurun_out/${tag}_ncu_abea.log 2>&1
tail -3 gpurun_out/${tag}_ncu_abea.log
