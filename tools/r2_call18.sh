#!/bin/bash
# Round 2, GPU call 18: one-block-per-row norm / RMSNorm-backward kernels, 2-MUFU GELU value + gradient in the gated-GELU backward,
# dx prefetch + 128-row blocks in the LoRA dx kernel -- against the previous commit's library (_prev) on the same box.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -4 ) > $O/c18_pytest.log 2>&1
tail -3 $O/c18_pytest.log
for v in "" _prev "" _prev; do
  ( MRB_LIB_VARIANT=$v timeout 300 python tools/t5_phase_bench.py ) > $O/c18_t5_phases$v.log 2>&1
  echo "variant [$v]"; tail -1 $O/c18_t5_phases$v.log
done
for v in "" _prev "" _prev; do
  ( MRB_LIB_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c18_bench$v.json 2> $O/c18_bench$v.err
  python -c "
import json; j=json.load(open('$O/c18_bench$v.json')); print('bench [$v]', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"
done
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > $O/c18_pytest_all.log 2>&1
tail -3 $O/c18_pytest_all.log
