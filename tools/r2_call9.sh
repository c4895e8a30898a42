#!/bin/bash
# Round 2, GPU call 9: programmatic dependent launch re-measured with alternating runs on ONE box (call 1 compared runs taken at
# different clock states); side streams on / off; GEMM sweep in the current build.
set -u
O=gpurun_out
mkdir -p $O
for i in 1 2; do
  ( timeout 300 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c9_bench_default_$i.json 2> $O/c9_bench_default_$i.err
  ( MRB_LIB_VARIANT=_pdl timeout 300 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c9_bench_pdl_$i.json 2> $O/c9_bench_pdl_$i.err
done
( MRB_OVERLAP=0 timeout 300 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c9_bench_nooverlap.json 2> $O/c9_bench_nooverlap.err
for f in default_1 pdl_1 default_2 pdl_2 nooverlap; do python -c "
import json; j=json.load(open('$O/c9_bench_$f.json')); print('$f', round(j['ms_per_step'],2), j['clocks']['sm_mhz'])"; done
( MRB_LIB_VARIANT=_pdl timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 ) > $O/c9_pytest_pdl.log 2>&1
tail -3 $O/c9_pytest_pdl.log
( timeout 200 python tools/gemm_sweep.py default $O/c9_sweep.json ) > $O/c9_sweep.log 2>&1
cut -c1-200 $O/c9_sweep.log | head -40
