#!/bin/bash
# Round 2, GPU call 12: ViT attention with descriptors formed next to their use, T5 phases in-graph, launch list of the step after
# the issue-path change, ncu of the fc2 / proj GEMMs (N = 1408) and of the T5 attention kernels.
set -u
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -5 ) > $O/c12_pytest_attn.log 2>&1
tail -2 $O/c12_pytest_attn.log
( timeout 200 python tools/attn_bench.py "" tc ) > $O/c12_attn_bench.log 2>&1
grep -v nobias $O/c12_attn_bench.log | cut -c1-120
( timeout 300 python tools/t5_phase_bench.py $O/c12_t5_phases.json ) > $O/c12_t5_phases.log 2>&1
tail -5 $O/c12_t5_phases.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/c12_launches.csv python tools/profile_one_step.py ) > $O/c12_ncu_list.log 2>&1
python tools/summarize_launches.py $O/c12_launches.csv $O/c12_launch_summary.csv; head -30 $O/c12_launch_summary.csv | cut -c1-120
gzip -f $O/c12_launches.csv
( MRB_DIAG_SHAPES=vit_fc2,vit_proj timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:gemm2 -o $O/c12_ncu_gemm_n1408 -f python tools/gemm_diag.py --ncu ) > $O/c12_ncu_gemm.log 2>&1
tail -2 $O/c12_ncu_gemm.log
( timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_.*tc -s 3 -c 3 -o $O/c12_ncu_attn_t5 -f python tools/attn_one.py ) > $O/c12_ncu_attn.log 2>&1
tail -2 $O/c12_ncu_attn.log
