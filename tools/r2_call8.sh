#!/bin/bash
# Round 2, GPU call 8: MMA issue reorder in the T5 attention kernels (S_{j+1} before P_j V_j; S^T_{t+1} before the accumulating MMAs).
set -u
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -30 ) > $O/c8_pytest_attn.log 2>&1
tail -4 $O/c8_pytest_attn.log
( MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py "" tc ) > $O/c8_attn_bench.log 2>&1
grep -v nobias $O/c8_attn_bench.log | cut -c1-120
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $O/c8_pytest.log 2>&1
tail -4 $O/c8_pytest.log
( timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c8_bench.json 2> $O/c8_bench.err
cut -c1-250 $O/c8_bench.json; tail -2 $O/c8_bench.err
