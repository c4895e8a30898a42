#!/bin/bash
# Round 2, GPU call 4: ViT kernel with the CLS-row scores moved to the softmax threads; exact delta for few-query-row backward.
set -u
O=gpurun_out
mkdir -p $O
( timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention_vit" 2>&1 | tail -30 ) > $O/c4_pytest_vit.log 2>&1
tail -5 $O/c4_pytest_vit.log
( timeout 120 python tools/attn_bench.py vit tc ) > $O/c4_attn_bench_vit.log 2>&1
cat $O/c4_attn_bench_vit.log | cut -c1-120
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/c4_pytest.log 2>&1
tail -6 $O/c4_pytest.log
( timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c4_bench.json 2> $O/c4_bench.err
cut -c1-250 $O/c4_bench.json; tail -2 $O/c4_bench.err
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_vit -s 1 -c 1 -o $O/c4_ncu_attn_vit -f python tools/attn_bench.py vit tc ) > $O/c4_ncu_attn_vit.log 2>&1
tail -2 $O/c4_ncu_attn_vit.log
