#!/bin/bash
# Round 2, GPU call 22: round-end evidence on one B200 -- full GPU suite, bench line with all three baselines, reference arm,
# launch list of the graph-replayed step, ncu --set full of the fc1 GEMM and of the T5 attention kernels, T5 phases, micro-benches.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $O/c22_pytest.log 2>&1
tail -3 $O/c22_pytest.log
( timeout 900 python bench.py --steps 10 --warmup 4 ) > $O/c22_bench.json 2> $O/c22_bench.err
cut -c1-400 $O/c22_bench.json; tail -2 $O/c22_bench.err
( timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/c22_bench_reference.json 2> $O/c22_bench_reference.err
cut -c1-300 $O/c22_bench_reference.json
( timeout 300 python tools/t5_phase_bench.py $O/c22_t5_phases.json ) > $O/c22_t5_phases.log 2>&1
tail -1 $O/c22_t5_phases.log
( MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py "" tc ) > $O/c22_attn_bench.log 2>&1
grep -v nobias $O/c22_attn_bench.log | cut -c1-120
( timeout 300 python tools/gemm_diag.py $O/c22_diag.json ) > $O/c22_diag.log 2>&1
grep -h "'name'" $O/c22_diag.log | cut -c1-200
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c22_launches.csv python tools/profile_one_step.py ) > $O/c22_ncu_list.log 2>&1
python tools/summarize_launches.py $O/c22_launches.csv $O/c22_launch_summary.csv; head -24 $O/c22_launch_summary.csv | cut -c1-100
gzip -f $O/c22_launches.csv
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 2 -c 1 -o $O/c22_ncu_fc1 -f python tools/gemm_one.py ) > $O/c22_ncu_fc1.log 2>&1
ncu -i $O/c22_ncu_fc1.ncu-rep --page raw --csv > $O/c22_ncu_fc1_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/c22_ncu_fc1_raw.csv $O/c22_ncu_fc1.csv | head -20
( MRB_ATTN_ONE_DROP=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_.*tc -s 3 -c 3 -o $O/c22_ncu_attn_t5 -f python tools/attn_one.py ) > $O/c22_ncu_attn.log 2>&1
tail -1 $O/c22_ncu_attn.log
