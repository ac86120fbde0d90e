#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02y_abea.txt; : > $O
timeout 600 python -m pytest tests/test_gpu_abea.py tests/test_gpu_prep.py -q 2>&1 | tail -3 >> $O
for v in old new; do
  echo "== abea $v" >> $O
  NPH_LIB_PATH=$PWD/nanopolish_b200/csrc/build/variants/libnph_abea_$v.so timeout 300 python scripts/quick_abea.py 2368 8000 2>&1 | tail -3 >> $O
done
cat $O
