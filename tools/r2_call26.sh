#!/bin/bash
# Round 2, GPU call 26: decoder-chain latency work -- split-K reduce folded into the GEMM launch (last-arriving CTA), block-per-row
# norms for decoder-sized inputs, wide small_down, 16-warp exact delta, 8-warp Q-Former cross-attention core.  Tests first, then
# the decoder chain in-graph under each switch, then the whole step old / new alternating on this box.
set -u
O=gpurun_out
mkdir -p $O
OLD="MRB_SPLITK_FUSED=0 MRB_NORM_ROW_SMALL=0 MRB_SMALL_DOWN_WIDE=0 MRB_DELTA_WIDE=0 MRB_XQ_WARPS=4"
( timeout 900 python -m pytest tests/test_splitk_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -8 ) > $O/c26_pytest.log 2>&1
tail -4 $O/c26_pytest.log
for v in "new:" "old:$OLD" "nofuse:MRB_SPLITK_FUSED=0" "nonorm:MRB_NORM_ROW_SMALL=0" "nodown:MRB_SMALL_DOWN_WIDE=0" "nodelta:MRB_DELTA_WIDE=0" "new2:"; do
  n=${v%%:*}; e=${v#*:}
  ( env $e MRB_T5_PHASES=dec_chain timeout 200 python tools/t5_phase_bench.py $O/c26_phase_$n.json ) > $O/c26_phase_$n.log 2>&1
  echo "dec_chain $n: $(tail -1 $O/c26_phase_$n.log | cut -c1-120)"
done
for v in old new old new; do
  e=""; [ $v = old ] && e="$OLD"
  ( env $e timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c26_bench_$v.json 2> $O/c26_bench_$v.err
  python -c "
import json; j=json.load(open('$O/c26_bench_$v.json')); print('bench $v', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3), round(j['qformer_xattn']['ms_per_step'],3), round(j['qformer_xattn']['frac'],3), j['gpu_launches'], j['loss'])"
done
