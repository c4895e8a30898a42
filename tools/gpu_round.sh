#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 400 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
( timeout 200 python tools/gemm_sweep.py default ) 2>&1 | grep "dec_\|lm_head\|down32" > $O/small.log
( MRB_GEMM_SMALL_A=0 timeout 200 python tools/gemm_sweep.py default ) 2>&1 | grep "dec_\|lm_head" | sed 's/^default/small_a_off/' >> $O/small.log
cat $O/small.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > $O/bench.json 2> $O/bench.err
cat $O/bench.json | cut -c1-300; tail -3 $O/bench.err
