#!/bin/bash
# Round 2, GPU call 16: tcgen05.ld of the next 16 columns in flight under the elementwise work of the current 16 (T5 attention
# forward softmax and backward dS loop).
set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -5 ) > $O/c16_pytest_attn.log 2>&1
tail -2 $O/c16_pytest_attn.log
( MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py "" tc ) > $O/c16_attn_bench.log 2>&1
grep -v "nobias" $O/c16_attn_bench.log | cut -c1-120
( timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c16_bench.json 2> $O/c16_bench.err
python -c "
import json; j=json.load(open('$O/c16_bench.json')); print('bench', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"
( timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_.*tc -s 3 -c 3 -o $O/c16_ncu_attn_t5 -f python tools/attn_one.py ) > $O/c16_ncu_attn.log 2>&1
tail -1 $O/c16_ncu_attn.log
