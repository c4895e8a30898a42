#!/bin/bash
# One GPU-box visit: parity tests, step breakdown (graph + eager), GEMM sweeps per kernel variant, bench line,
# ncu --set full of the fc1 GEMM.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
( timeout 600 python tools/step_breakdown.py $O/step_breakdown.json 2>&1 | tail -5 ) > $O/breakdown.log 2>&1
python tools/print_breakdown.py 2>&1 | head -20
( timeout 300 python tools/gemm_sweep.py default $O/sweep_default.json ) > $O/sweep_default.log 2>&1
( MRB_GEMM2_EPI=plain timeout 300 python tools/gemm_sweep.py epi_plain $O/sweep_epi_plain.json ) > $O/sweep_epi_plain.log 2>&1
( MRB_GEMM2_TAIL=0 timeout 300 python tools/gemm_sweep.py tail_off $O/sweep_tail_off.json ) > $O/sweep_tail_off.log 2>&1
cat $O/sweep_default.log $O/sweep_epi_plain.log $O/sweep_tail_off.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > $O/bench.json 2> $O/bench.err
cat $O/bench.json; tail -3 $O/bench.err
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 2 -c 1 -o $O/ncu_fc1 -f python tools/gemm_one.py ) > $O/ncu_fc1.log 2>&1
tail -3 $O/ncu_fc1.log
