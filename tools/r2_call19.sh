#!/bin/bash
# Round 2, GPU call 19: coalesced delta kernel, full GPU suite (incl. the full-depth gate), one-block-per-row forward norms on / off.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $O/c19_pytest_all.log 2>&1
tail -3 $O/c19_pytest_all.log
( MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py t5enc tc ) > $O/c19_attn_bench.log 2>&1
cat $O/c19_attn_bench.log | cut -c1-120
for v in 1 0 1 0; do
  ( MRB_NORM_ROW=$v timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c19_bench_$v.json 2> $O/c19_bench_$v.err
  python -c "
import json; j=json.load(open('$O/c19_bench_$v.json')); print('bench norm_row=$v', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"
done
