#!/bin/bash
# Round 2, GPU call 25: the two frame halves of the ViT on two streams (MRB_VIT_SPLIT=1) vs one stream.
set -u
O=gpurun_out
mkdir -p $O
for v in 0 1 0 1; do
  ( MRB_VIT_SPLIT=$v timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c25_bench_$v.json 2> $O/c25_bench_$v.err
  python -c "
import json; j=json.load(open('$O/c25_bench_$v.json')); print('bench vit_split=$v', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3), j['loss'])"
done
( MRB_VIT_SPLIT=1 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_full_size_gpu.py -m gpu -q -x 2>&1 | tail -3 ) > $O/c25_pytest_split.log 2>&1
tail -2 $O/c25_pytest_split.log
