#!/bin/bash
# Round 2, GPU call 10: what holds the 2-CTA GEMM at ~58 % tensor-pipe activity?  This library vs cuBLAS on the same operands
# (alternating), diagnostic builds (main loop alone / MMA issue alone / 4-stage ring), the direct-store epilogue with a 7-stage
# ring, and one `ncu --set full` capture of both on two shapes (source-level stall samples of ours, configuration of theirs).
set -u
O=gpurun_out
mkdir -p $O
( timeout 300 python tools/gemm_diag.py $O/c10_diag_default.json ) > $O/c10_diag_default.log 2>&1
for v in _direct _noepi _notma _st4; do
  ( MRB_LIB_VARIANT=$v MRB_DIAG_SHAPES=vit_qkv,vit_proj,vit_fc1,vit_fc2,t5_qkv,t5_o,t5_wo timeout 300 python tools/gemm_diag.py $O/c10_diag$v.json ) > $O/c10_diag$v.log 2>&1
done
grep -h "'name'" $O/c10_diag_*.log | cut -c1-230
( MRB_LIB_VARIANT=_direct timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or patch" 2>&1 | tail -5 ) > $O/c10_pytest_direct.log 2>&1
tail -3 $O/c10_pytest_direct.log
( timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -o $O/c10_ncu_gemm -f python tools/gemm_diag.py --ncu ) > $O/c10_ncu_gemm.log 2>&1
tail -3 $O/c10_ncu_gemm.log
( MRB_LIB_VARIANT=_direct timeout 300 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c10_bench_direct.json 2> $O/c10_bench_direct.err
( timeout 300 python bench.py --steps 8 --warmup 3 --no-eager --no-cpu-baseline ) > $O/c10_bench_default.json 2> $O/c10_bench_default.err
for f in direct default; do python -c "
import json; j=json.load(open('$O/c10_bench_$f.json')); print('$f', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], j['roofline']['frac'])"; done
