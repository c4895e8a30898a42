#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
./tools/probe/cluster_probe > $O/cluster_probe.log 2>&1; cat $O/cluster_probe.log
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
( timeout 300 python tools/attn_bench.py vit tc ) > $O/attn_bench.log 2>&1
( MRB_ATTN_G1=0 timeout 300 python tools/attn_bench.py vit tc ) >> $O/attn_bench.log 2>&1
cat $O/attn_bench.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > $O/bench.json 2> $O/bench.err
cat $O/bench.json; tail -3 $O/bench.err
