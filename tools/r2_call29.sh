#!/bin/bash
# Round 2, GPU call 29: round-end evidence of the third session on one B200 -- full GPU suite, bench line with all three baselines,
# reference arm, T5 phases, launch list of the graph-replayed step, memcheck over a 2-layer step.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $O/c29_pytest.log 2>&1
tail -3 $O/c29_pytest.log
( timeout 900 python bench.py --steps 10 --warmup 4 ) > $O/c29_bench.json 2> $O/c29_bench.err
cut -c1-400 $O/c29_bench.json; tail -2 $O/c29_bench.err
( timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/c29_bench_reference.json 2> $O/c29_bench_reference.err
cut -c1-300 $O/c29_bench_reference.json
( timeout 300 python tools/t5_phase_bench.py $O/c29_t5_phases.json ) > $O/c29_t5_phases.log 2>&1
tail -1 $O/c29_t5_phases.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c29_launches.csv python tools/profile_one_step.py ) > $O/c29_ncu_list.log 2>&1
python tools/summarize_launches.py $O/c29_launches.csv $O/c29_launch_summary.csv; head -24 $O/c29_launch_summary.csv | cut -c1-100
gzip -f $O/c29_launches.csv
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step.py 60 > $O/c29_sanitize_memcheck.log 2>&1
grep -H "ERROR SUMMARY\|loss" $O/c29_sanitize_memcheck.log | tail -4
