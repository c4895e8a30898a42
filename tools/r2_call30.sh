#!/bin/bash
# Round 2, GPU call 30: full GPU suite after the SM cap became a per-thread setting applied inside T5Engine.side_block (call 29: three
# whole-model tests failed once a process had built more than eight engines -- the per-stream table of the first version was full);
# bench line of the final code; the other BASELINE.json configs on one B200 with the third session's defaults.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 ) > $O/c30_pytest.log 2>&1
tail -3 $O/c30_pytest.log
( timeout 900 python bench.py --steps 10 --warmup 4 ) > $O/c30_bench.json 2> $O/c30_bench.err
cut -c1-200 $O/c30_bench.json; tail -1 $O/c30_bench.err
for c in charades anet generate; do
  ( timeout 500 python bench.py --config $c --steps 6 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c30_bench_$c.json 2> $O/c30_bench_$c.err
  python -c "
import json; j=json.load(open('$O/c30_bench_$c.json')); print('$c', round(j['ms_per_step'],2), 'ms', round(j['value'],2), j['unit'], 'e2e', round(j['e2e']['value'],2), 'roofline', round(j['roofline']['frac'],3), j['clocks']['sm_mhz'])" || tail -3 $O/c30_bench_$c.err
done
