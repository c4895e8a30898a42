#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for V in "" _hint _hint2; do
  export MRB_LIB_VARIANT=$V
  echo "== variant '$V'" >> $O/variants.log
  ( timeout 200 python tools/attn_bench.py "" tc ) >> $O/variants.log 2>&1
  ( timeout 200 python tools/gemm_sweep.py v$V ) 2>&1 | grep "vit_\|t5_" >> $O/variants.log
  ( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 ) | grep "timed region" >> $O/variants.log
done
cat $O/variants.log
