#!/bin/bash
# Round 2, GPU call 28: fused residual-stream passes of the T5 train step (dropout-add + next RMSNorm; RMSNorm backward + next masked
# gradient operand: MRB_T5_FUSE_NORM), defaults after call 27 (side-stream SM cap 132, two ViT streams); ncu --set full of the
# Q-Former cross-attention path's two kernels (K/V projection GEMM, attention core).
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_model_gpu.py tests/test_full_size_gpu.py -m gpu -q -x 2>&1 | tail -6 ) > $O/c28_pytest.log 2>&1
tail -3 $O/c28_pytest.log
for v in 0 1 0 1; do
  ( MRB_T5_FUSE_NORM=$v timeout 300 python tools/t5_phase_bench.py $O/c28_phases_fuse$v.json ) > $O/c28_phases_fuse$v.log 2>&1
  echo "phases fuse=$v: $(tail -1 $O/c28_phases_fuse$v.log | cut -c1-160)"
done
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_xq|gemm2" -s 7 -c 2 -o $O/c28_ncu_xattn -f python tools/xattn_one.py ) > $O/c28_ncu_xattn.log 2>&1
tail -2 $O/c28_ncu_xattn.log
python tools/ncu_kernels.py $O/c28_ncu_xattn.ncu-rep $O/c28_ncu_xattn.md > /dev/null 2>&1; grep -E "^## |time_duration|tensor_cycles|dram__b|dram__thr|warps_active" $O/c28_ncu_xattn.md | cut -c1-150
( timeout 120 python tools/xattn_one.py ) 2>&1 | tail -1
for v in 0 1 0 1; do
  ( MRB_T5_FUSE_NORM=$v timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c28_bench_fuse$v.json 2> $O/c28_bench_fuse$v.err
  python -c "
import json; j=json.load(open('$O/c28_bench_fuse$v.json')); x=j['qformer_xattn']; print('bench fuse=$v', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3), 'xattn', round(x['ms_per_step'],3), round(x['frac'],3), j['gpu_launches'], j['loss'])"
done
