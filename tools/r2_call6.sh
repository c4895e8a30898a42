#!/bin/bash
# Round 2, GPU call 6: mma exact-delta, restructured LoRA wgrad / dx kernels, TMEM read probe.
set -u
O=gpurun_out
mkdir -p $O
( timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_model_gpu.py -m gpu -q 2>&1 | tail -30 ) > $O/c6_pytest_new.log 2>&1
tail -5 $O/c6_pytest_new.log
( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tmem_ld_probe tools/probe/tmem_ld_probe.cu 2>&1 | grep -v deprecated; timeout 60 /tmp/tmem_ld_probe ) > $O/c6_tmem_ld_probe.log 2>&1
cat $O/c6_tmem_ld_probe.log
( timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c6_bench.json 2> $O/c6_bench.err
cut -c1-250 $O/c6_bench.json; tail -2 $O/c6_bench.err
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c6_launches.csv python tools/profile_one_step.py ) > $O/c6_ncu_list.log 2>&1
python tools/summarize_launches.py $O/c6_launches.csv $O/c6_launch_summary.csv > /dev/null 2>&1
head -45 $O/c6_launch_summary.csv | cut -c1-150
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $O/c6_pytest.log 2>&1
tail -4 $O/c6_pytest.log
