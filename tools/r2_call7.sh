#!/bin/bash
# Round 2, GPU call 7: half-tile pipelined attention backward, deterministic exact delta.
set -u
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -30 ) > $O/c7_pytest_attn.log 2>&1
tail -6 $O/c7_pytest_attn.log
( MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py "" tc ) > $O/c7_attn_bench.log 2>&1
cat $O/c7_attn_bench.log | cut -c1-120
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $O/c7_pytest.log 2>&1
tail -5 $O/c7_pytest.log
( timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c7_bench.json 2> $O/c7_bench.err
cut -c1-250 $O/c7_bench.json; tail -2 $O/c7_bench.err
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_.*tc -s 4 -c 3 -o $O/c7_ncu_attn_t5 -f python tools/attn_one.py ) > $O/c7_ncu_attn_t5.log 2>&1
tail -2 $O/c7_ncu_attn_t5.log
