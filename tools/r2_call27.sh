#!/bin/bash
# Round 2, GPU call 27: SM cap of the decoder's side-stream GEMMs (MRB_SIDE_SMS), the defaults after call 26 (two-pass split-K
# reduce, four-warp cross-attention core with K / V in two cp.async groups), the cross-attention path timed back to back in a graph.
set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests/test_splitk_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -5 ) > $O/c27_pytest.log 2>&1
tail -2 $O/c27_pytest.log
best=0; bestms=999999
for sms in 0 132 116 100 84; do
  ( MRB_SIDE_SMS=$sms MRB_T5_PHASES=dec_chain timeout 200 python tools/t5_phase_bench.py $O/c27_phase_$sms.json ) > $O/c27_phase_$sms.log 2>&1
  ms=$(python -c "import json; print(json.load(open('$O/c27_phase_$sms.json'))['dec_chain'])" 2>/dev/null || echo 999999)
  echo "dec_chain side_sms=$sms: $ms ms"
  if python -c "import sys; sys.exit(0 if float('$ms') < float('$bestms') else 1)"; then best=$sms; bestms=$ms; fi
done
echo "best side_sms=$best ($bestms ms)"
echo $best > $O/c27_best_side_sms.txt
for v in base side base side split; do
  e=""; [ $v = side ] && e="MRB_SIDE_SMS=$best"; [ $v = split ] && e="MRB_SIDE_SMS=$best MRB_VIT_SPLIT=1"
  ( env $e timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c27_bench_$v.json 2> $O/c27_bench_$v.err
  python -c "
import json; j=json.load(open('$O/c27_bench_$v.json')); x=j['qformer_xattn']; print('bench $v', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3), 'xattn graph', round(x['ms_per_step'],3), round(x['frac'],3), 'eager', round(x['ms_eager_events'],3), j['loss'])"
done
