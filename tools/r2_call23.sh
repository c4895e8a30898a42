#!/bin/bash
# Round 2, GPU call 23: the attention backward measured 0.656 / 0.875 ms with 196 KB of dynamic shared memory (64 KB of it unused after
# the operands moved to tensor memory) and 0.707 / 0.927 ms with 130 KB: padding vs a deeper streamed-tile ring vs the lean layout.
set -u
O=gpurun_out
mkdir -p $O
for v in "" _pad _st6 "" _pad _st6; do
  ( MRB_LIB_VARIANT=$v MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py t5enc tc ) > $O/c23_attn_bench$v.log 2>&1
  echo "variant [$v]"; grep bwd $O/c23_attn_bench$v.log | cut -c1-120
done
( MRB_LIB_VARIANT=_st6 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -3 ) > $O/c23_pytest_st6.log 2>&1
tail -2 $O/c23_pytest_st6.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step.py 60 > $O/c23_sanitize_memcheck.log 2>&1
grep -H "ERROR SUMMARY\|loss" $O/c23_sanitize_memcheck.log | tail -4
