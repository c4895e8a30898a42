#!/bin/bash
# Round 2, GPU call 2: the un-gated suite (dropout / QA / split-K now default), the full-depth parity test, smoke(), the new
# bench line (train-mode dropout default, eager arm, CPU sample), the other configs, launch list of one dropout step.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > $O/c2_pytest.log 2>&1
tail -5 $O/c2_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/c2_smoke.log 2>&1
tail -2 $O/c2_smoke.log
( timeout 900 python bench.py --steps 8 --warmup 3 ) > $O/c2_bench.json 2> $O/c2_bench.err
cut -c1-250 $O/c2_bench.json; tail -3 $O/c2_bench.err
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/c2_bench_reference.json 2> $O/c2_bench_reference.err
cut -c1-200 $O/c2_bench_reference.json
for c in charades anet generate; do
  ( timeout 600 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline ) > $O/c2_bench_$c.json 2> $O/c2_bench_$c.err
  cut -c1-250 $O/c2_bench_$c.json; tail -2 $O/c2_bench_$c.err
done
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c2_launches_dropout.csv python tools/profile_one_step.py ) > $O/c2_ncu_list.log 2>&1
python tools/summarize_launches.py $O/c2_launches_dropout.csv $O/c2_launch_summary_dropout.csv > /dev/null 2>&1
head -45 $O/c2_launch_summary_dropout.csv | cut -c1-150
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 2 -c 1 -o $O/c2_ncu_attn_t5_fwd -f python tools/attn_one.py ) > $O/c2_ncu_attn_t5_fwd.log 2>&1
tail -2 $O/c2_ncu_attn_t5_fwd.log
