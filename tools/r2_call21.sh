#!/bin/bash
# Round 2, GPU call 21: P^T / dS^T operands of the accumulating MMAs in tensor memory (TS-form tcgen05.mma) in the T5 attention backward.
set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -5 ) > $O/c21_pytest_attn.log 2>&1
tail -2 $O/c21_pytest_attn.log
( MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py "" tc ) > $O/c21_attn_bench.log 2>&1
grep -v "nobias\|^vit" $O/c21_attn_bench.log | cut -c1-120
( MRB_ATTN_BENCH_DROP=1 timeout 300 python tools/attn_bench.py cross ) > $O/c21_cross.log 2>&1
grep "tc " $O/c21_cross.log | cut -c1-120
( timeout 300 python tools/t5_phase_bench.py ) > $O/c21_t5_phases.log 2>&1
tail -1 $O/c21_t5_phases.log
( timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c21_bench.json 2> $O/c21_bench.err
python -c "
import json; j=json.load(open('$O/c21_bench.json')); print('bench', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"
