#!/bin/bash
# Round-end evidence run on one B200: parity tests, bench line (+ CPU baseline, + reference arm), step breakdown,
# GEMM / attention micro-benches, ncu launch list of one graph-replayed step, ncu --set full of the fc1 GEMM.
set -u
O=gpurun_out
mkdir -p $O
( timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/pytest.log 2>&1
tail -3 $O/pytest.log
( timeout 400 python bench.py --steps 8 --warmup 3 ) > $O/bench.json 2> $O/bench.err
cat $O/bench.json | cut -c1-300; tail -2 $O/bench.err
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err
cat $O/bench_reference.json | cut -c1-300
( timeout 300 python tools/step_breakdown.py $O/step_breakdown.json 2>&1 | tail -3 ) > $O/breakdown.log 2>&1
( timeout 200 python tools/gemm_sweep.py default $O/sweep_default.json ) > $O/sweep_default.log 2>&1
( timeout 200 python tools/attn_bench.py "" tc ) > $O/attn_bench.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_one_step.py ) > $O/ncu_list.log 2>&1
python tools/summarize_launches.py $O/launches.csv $O/launch_summary.csv
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 2 -c 1 -o $O/ncu_fc1 -f python tools/gemm_one.py ) > $O/ncu_fc1.log 2>&1
ncu -i $O/ncu_fc1.ncu-rep --page raw --csv > $O/ncu_fc1_raw.csv 2>/dev/null
tail -2 $O/ncu_fc1.log
