#!/bin/bash
# Round 2, GPU call 35: the step as two graphs (front: ViT .. t5_proj, launched before the host builds the prompt table; back: T5)
# -- MRB_SPLIT_GRAPH=1 -- and no second, eager LoRA re-pack in front of a graphed step: model tests both ways, e2e A/B.
set -u
O=gpurun_out
mkdir -p $O
( MRB_SPLIT_GRAPH=1 timeout 400 python -m pytest tests/test_model_gpu.py tests/test_full_size_gpu.py -m gpu -q --tb=short 2>&1 | tail -12 ) > $O/c35_pytest_split.log 2>&1
tail -2 $O/c35_pytest_split.log
( timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short 2>&1 | tail -12 ) > $O/c35_pytest_default.log 2>&1
tail -2 $O/c35_pytest_default.log
for v in 0 1; do
  ( MRB_SPLIT_GRAPH=$v timeout 300 python bench.py --steps 8 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c35_bench_$v.json 2> $O/c35_bench_$v.err
  python -c "
import json; j=json.load(open('$O/c35_bench_$v.json')); print('split_graph=$v value', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['ms_per_step'],2), 'u8', round(j['e2e_uint8_frames']['ms_per_step'],2), j['clocks']['sm_mhz'], j['gpu_launches'], j['loss'])" || tail -3 $O/c35_bench_$v.err
done
