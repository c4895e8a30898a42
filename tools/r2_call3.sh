#!/bin/bash
# Round 2, GPU call 3: the persistent ViT attention kernel (unit test first, under a short timeout: new barrier protocol), its
# timing next to the round-1 kernel, the fixed tests, the all-gradients dump of the full-depth parity test.
set -u
O=gpurun_out
mkdir -p $O
( timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention_vit" 2>&1 | tail -40 ) > $O/c3_pytest_vit.log 2>&1
tail -25 $O/c3_pytest_vit.log
( timeout 120 python tools/attn_bench.py vit tc ) > $O/c3_attn_bench_vit.log 2>&1
cat $O/c3_attn_bench_vit.log | cut -c1-120
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/c3_pytest.log 2>&1
tail -6 $O/c3_pytest.log
( timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c3_bench.json 2> $O/c3_bench.err
cut -c1-250 $O/c3_bench.json; tail -2 $O/c3_bench.err
( MRB_ATTN_VIT=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c3_bench_oldvit.json 2> $O/c3_bench_oldvit.err
cut -c1-250 $O/c3_bench_oldvit.json
