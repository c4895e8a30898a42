#!/bin/bash
# Round 2, GPU call 5: tensor-core LoRA-dropout kernels, ViT kernel (dtype templated), tie-aware tests.
set -u
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "attention_vit or lora_dropout" 2>&1 | tail -30 ) > $O/c5_pytest_new.log 2>&1
tail -5 $O/c5_pytest_new.log
( timeout 120 python tools/attn_bench.py vit tc ) > $O/c5_attn_bench_vit.log 2>&1
cat $O/c5_attn_bench_vit.log | cut -c1-120
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/c5_pytest.log 2>&1
tail -6 $O/c5_pytest.log
( timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c5_bench.json 2> $O/c5_bench.err
cut -c1-250 $O/c5_bench.json; tail -2 $O/c5_bench.err
( MRB_LORA_DROP_MMA=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c5_bench_oldlora.json 2> $O/c5_bench_oldlora.err
cut -c1-250 $O/c5_bench_oldlora.json
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c5_launches.csv python tools/profile_one_step.py ) > $O/c5_ncu_list.log 2>&1
python tools/summarize_launches.py $O/c5_launches.csv $O/c5_launch_summary.csv > /dev/null 2>&1
head -40 $O/c5_launch_summary.csv | cut -c1-150
