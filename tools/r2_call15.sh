#!/bin/bash
# Round 2, GPU call 15: mbarrier.try_wait with a suspend-time hint (-DMRB_WAIT_HINT_NS=10000000, the value CUTLASS passes): waiting
# warps sleep on the barrier instead of spinning through try_wait / branch / spin-count instructions that take issue slots from
# the working warps of their scheduler (ncu instruction mix of the attention forward: ~35 % integer / branch / clock-read work).
set -u
O=gpurun_out
mkdir -p $O
for v in "" _hint "" _hint; do
  ( MRB_LIB_VARIANT=$v MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py "" tc ) > $O/c15_attn_bench$v.log 2>&1
  echo "variant [$v]"; grep -v "nobias\|4017" $O/c15_attn_bench$v.log | cut -c1-120
done
for v in "" _hint; do
  ( MRB_LIB_VARIANT=$v MRB_DIAG_SHAPES=vit_qkv,vit_proj,vit_fc1,vit_fc2,t5_qkv,t5_o,t5_wo timeout 300 python tools/gemm_diag.py $O/c15_diag$v.json ) > $O/c15_diag$v.log 2>&1
  grep -h "'name'" $O/c15_diag$v.log | cut -c1-200
done
( MRB_LIB_VARIANT=_hint timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x 2>&1 | tail -3 ) > $O/c15_pytest_hint.log 2>&1
tail -2 $O/c15_pytest_hint.log
for v in "" _hint "" _hint; do
  ( MRB_LIB_VARIANT=$v timeout 300 python tools/t5_phase_bench.py ) > $O/c15_t5_phases$v.log 2>&1
  echo "variant [$v]"; tail -1 $O/c15_t5_phases$v.log
done
for v in "" _hint "" _hint; do
  ( MRB_LIB_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c15_bench$v.json 2> $O/c15_bench$v.err
  python -c "
import json; j=json.load(open('$O/c15_bench$v.json')); print('bench [$v]', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"
done
