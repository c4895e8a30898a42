#!/bin/bash
# Round 2, first GPU call: everything round 1 left unverified on hardware, in dependency order, plus the ncu --set full
# captures of the attention kernels the verdict asked for.  Outputs -> gpurun_out/c1_*.
set -u
O=gpurun_out
mkdir -p $O
( timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/c1_pytest.log 2>&1
tail -3 $O/c1_pytest.log
( MRB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q 2>&1 | tail -40 ) > $O/c1_pytest_splitk.log 2>&1
tail -5 $O/c1_pytest_splitk.log
( MRB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_dropout_gpu.py tests/test_qa_gpu.py -m gpu -q 2>&1 | tail -150 ) > $O/c1_pytest_dropout.log 2>&1
tail -8 $O/c1_pytest_dropout.log
( timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline ) > $O/c1_bench.json 2> $O/c1_bench.err
( timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --train-dropout ) > $O/c1_bench_dropout.json 2> $O/c1_bench_dropout.err
( MRB_GEMM_SPLITK=1 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline ) > $O/c1_bench_splitk.json 2> $O/c1_bench_splitk.err
( MRB_LIB_VARIANT=_pdl timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline ) > $O/c1_bench_pdl.json 2> $O/c1_bench_pdl.err
cut -c1-220 $O/c1_bench.json $O/c1_bench_dropout.json $O/c1_bench_splitk.json $O/c1_bench_pdl.json
tail -2 $O/c1_bench_dropout.err $O/c1_bench_pdl.err
( MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py "" tc ) > $O/c1_attn_bench.log 2>&1
cat $O/c1_attn_bench.log | cut -c1-120
( timeout 200 python tools/gemm_sweep.py default $O/c1_sweep_default.json ) > $O/c1_sweep_default.log 2>&1
( MRB_GEMM_SPLITK=1 timeout 200 python tools/gemm_sweep.py splitk $O/c1_sweep_splitk.json ) > $O/c1_sweep_splitk.log 2>&1
grep -h "dec_\|down32\|lm_head" $O/c1_sweep_default.log $O/c1_sweep_splitk.log | cut -c1-220
# ncu --set full of the attention kernels (T5 encoder fwd <64,2,1>, bwd dKV + dQ; ViT fwd <96,1,0>)
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_.*tc -s 4 -c 3 -o $O/c1_ncu_attn_t5 -f python tools/attn_one.py ) > $O/c1_ncu_attn_t5.log 2>&1
tail -2 $O/c1_ncu_attn_t5.log
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 1 -c 1 -o $O/c1_ncu_attn_vit -f python tools/attn_bench.py vit tc ) > $O/c1_ncu_attn_vit.log 2>&1
tail -2 $O/c1_ncu_attn_vit.log
( MRB_LIB_VARIANT=_pdl timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/c1_pytest_pdl.log 2>&1
tail -3 $O/c1_pytest_pdl.log
( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_feed_probe tools/probe/tma_feed_probe.cu -lcuda 2>&1 | grep -v deprecated; timeout 120 /tmp/tma_feed_probe ) > $O/c1_tma_feed_probe.log 2>&1
tail -45 $O/c1_tma_feed_probe.log
