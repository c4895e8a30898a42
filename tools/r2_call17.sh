#!/bin/bash
# Round 2, GPU call 17: one-block-per-row RMSNorm backward for the encoder-sized calls (A/B in-graph), ncu --set full of the
# HBM-bound / mask-bound train-mode kernels of the encoder layers inside one graph-replayed step.
set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "norm" 2>&1 | tail -3 ) > $O/c17_pytest_norm.log 2>&1
tail -2 $O/c17_pytest_norm.log
for v in 1 0 1 0; do
  ( MRB_RMSNORM_BWD_ROW=$v timeout 300 python tools/t5_phase_bench.py ) > $O/c17_t5_phases_$v.log 2>&1
  echo "row kernel $v"; tail -1 $O/c17_t5_phases_$v.log
done
( timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"lora_dx_drop|gated_gelu_bwd_drop|gated_gelu_fwd_drop|lora_down_drop|rmsnorm_bwd_row|dropout_add_kernel|lora_wgrad_drop" --launch-skip 700 -c 28 -o $O/c17_ncu_elt -f python tools/profile_one_step.py ) > $O/c17_ncu_elt.log 2>&1
tail -2 $O/c17_ncu_elt.log
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > $O/c17_pytest.log 2>&1
tail -3 $O/c17_pytest.log
