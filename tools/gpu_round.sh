#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
( timeout 300 python tools/gemm_sweep.py default $O/sweep_default.json ) > $O/sweep_default.log 2>&1
( MRB_GEMM2_EPI=generic timeout 300 python tools/gemm_sweep.py epi_generic $O/sweep_epi_generic.json ) > $O/sweep_epi_generic.log 2>&1
grep -h "vit_\|qf_\|t5_" $O/sweep_default.log $O/sweep_epi_generic.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > $O/bench.json 2> $O/bench.err
cat $O/bench.json; tail -3 $O/bench.err
( MRB_OVERLAP=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > $O/bench_nooverlap.json 2> $O/bench_nooverlap.err
cat $O/bench_nooverlap.json; tail -3 $O/bench_nooverlap.err
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 2 -c 1 -o $O/ncu_fc1 -f python tools/gemm_one.py ) > $O/ncu_fc1.log 2>&1
tail -3 $O/ncu_fc1.log
