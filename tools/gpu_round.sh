#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > $O/bench.json 2> $O/bench.err
cat $O/bench.json | cut -c1-300; tail -3 $O/bench.err
