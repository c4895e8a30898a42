#!/bin/bash
# Round 2, GPU call 24: 3-stage K/V ring in the T5 attention forward (202 KB of shared memory) vs 2 stages.
set -u
O=gpurun_out
mkdir -p $O
for v in "" _f3 "" _f3; do
  ( MRB_LIB_VARIANT=$v MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py t5enc tc ) > $O/c24_attn_bench$v.log 2>&1
  echo "variant [$v]"; grep -v bwd $O/c24_attn_bench$v.log | cut -c1-120
done
( MRB_LIB_VARIANT=_f3 MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py t5enc_4017 tc ) > $O/c24_attn_bench_4017_f3.log 2>&1
grep -v bwd $O/c24_attn_bench_4017_f3.log | cut -c1-120
( MRB_LIB_VARIANT=_f3 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -3 ) > $O/c24_pytest_f3.log 2>&1
tail -2 $O/c24_pytest_f3.log
