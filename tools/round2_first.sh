#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, in dependency order.
#   1. the default GPU suite (must still be green: gemm.cu host dispatch, ops.py and config.py changed since the last GPU run)
#   2. the experimental split-K tests (tests/test_experimental_gpu.py)
#   2b. the train-mode dropout path (tests/test_dropout_gpu.py: every kernel against the oracle's masks, engines, model, graphs),
#       then the bench line with it -- if green: make train_dropout the default and the bench workload
#   3. small-shape sweep without / with split-K, and the bench line without / with it
# If 2 is green and 3 shows the gain: flip ops.SPLITK's default, move the tests into tests/test_kernels_gpu.py.
set -u
O=gpurun_out
mkdir -p $O
( timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/pytest.log 2>&1
tail -3 $O/pytest.log
( MRB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q 2>&1 | tail -25 ) > $O/pytest_experimental.log 2>&1
tail -5 $O/pytest_experimental.log
( MRB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_dropout_gpu.py tests/test_qa_gpu.py -m gpu -q 2>&1 | tail -60 ) > $O/pytest_dropout.log 2>&1
tail -8 $O/pytest_dropout.log
( timeout 400 python bench.py --steps 8 --warmup 3 --train-dropout ) > $O/bench_dropout.json 2> $O/bench_dropout.err
cut -c1-200 $O/bench_dropout.json
# where the dropout step's extra time goes: launch list of one graph-replayed step with dropout on (compare with launch_summary_r01b.csv)
( MRB_TRAIN_DROPOUT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_dropout.csv python tools/profile_one_step.py ) > $O/ncu_list_dropout.log 2>&1
python tools/summarize_launches.py $O/launches_dropout.csv $O/launch_summary_dropout.csv
( MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py "" tc ) > $O/attn_bench_dropout.log 2>&1
grep -h "dropout\|bwd" $O/attn_bench_dropout.log | cut -c1-120
( timeout 200 python tools/gemm_sweep.py default $O/sweep_default.json ) > $O/sweep_default.log 2>&1
( MRB_GEMM_SPLITK=1 timeout 200 python tools/gemm_sweep.py splitk $O/sweep_splitk.json ) > $O/sweep_splitk.log 2>&1
grep -h "dec_\|down32\|lm_head" $O/sweep_default.log $O/sweep_splitk.log | cut -c1-220
( timeout 400 python bench.py --steps 8 --warmup 3 ) > $O/bench.json 2> $O/bench.err
( MRB_GEMM_SPLITK=1 timeout 400 python bench.py --steps 8 --warmup 3 ) > $O/bench_splitk.json 2> $O/bench_splitk.err
cut -c1-200 $O/bench.json $O/bench_splitk.json
( MRB_GEMM_SPLITK=1 timeout 300 python tools/run_configs.py $O/configs_splitk.json 2>&1 | tail -5 ) > $O/configs_splitk.log 2>&1
# what bounds the operand feed of one SM (decides: 64-row A boxes? multicast? more CTAs per problem?)
( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_feed_probe tools/probe/tma_feed_probe.cu -lcuda 2>&1 | grep -v deprecated; timeout 120 /tmp/tma_feed_probe ) > $O/tma_feed_probe.log 2>&1
tail -45 $O/tma_feed_probe.log
# programmatic dependent launch (build flag -DMRB_PDL: second library, default build untouched): the whole GPU suite and the
# bench line on the variant.  Adopt (make it the default build) only if the suite is green and the step gets shorter.
( python -m mr_blip_b200.build --variant _pdl -DMRB_PDL 2>&1 | tail -1 ) > $O/build_pdl.log 2>&1
( MRB_LIB_VARIANT=_pdl timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/pytest_pdl.log 2>&1
tail -3 $O/pytest_pdl.log
( MRB_LIB_VARIANT=_pdl timeout 400 python bench.py --steps 8 --warmup 3 ) > $O/bench_pdl.json 2> $O/bench_pdl.err
cut -c1-200 $O/bench_pdl.json
