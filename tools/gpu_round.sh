#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
( timeout 300 python tools/attn_bench.py "" tc ) > $O/attn_bench.log 2>&1
( MRB_ATTN_BWD_ENC=0 timeout 300 python tools/attn_bench.py t5enc tc ) >> $O/attn_bench.log 2>&1
cat $O/attn_bench.log
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 -o $O/ncu_attn -f python tools/attn_one.py ) > $O/ncu_attn.log 2>&1
tail -3 $O/ncu_attn.log
