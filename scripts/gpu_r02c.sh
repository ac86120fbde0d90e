#!/bin/bash
# A/B of K1's row update: each variant library on the scorereads and methylation-window workloads (resident kernel time)
mkdir -p gpurun_out
out=gpurun_out/r02c_k1_variants.txt; : > $out
for v in scalar_lea scalar_imad packed_all packed_all_lea packed_arith packed_lsum packed_all_c9; do
  for w in scorereads methylation; do
    NPH_LIB_PATH=$PWD/nanopolish_b200/csrc/build/variants/libnph_$v.so timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-call-methylation 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '$w', 'value=%.4g' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'], 'cells/s=%.4g' % d['roofline']['block_cells_per_sec_per_gpu'], 'e2e=%.4g' % d['e2e']['value'])" >> $out 2>&1
  done
done
cat $out
