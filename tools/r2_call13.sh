#!/bin/bash
# Round 2, GPU call 13: few-query cross-attention kernels (keys split over a cluster), one-row-per-block norms for decoder-sized
# inputs, ViT attention back on its round-2b issue path, programmatic dependent launch restricted to small grids (_pdls build).
set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x 2>&1 | tail -30 ) > $O/c13_pytest_kernels.log 2>&1
tail -4 $O/c13_pytest_kernels.log
( timeout 200 python tools/attn_bench.py vit tc ) > $O/c13_attn_bench.log 2>&1
cat $O/c13_attn_bench.log | cut -c1-120
( timeout 300 python tools/t5_phase_bench.py $O/c13_t5_phases.json ) > $O/c13_t5_phases.log 2>&1
tail -1 $O/c13_t5_phases.log
( MRB_ATTN_FQ=0 timeout 300 python tools/t5_phase_bench.py $O/c13_t5_phases_nofq.json ) > $O/c13_t5_phases_nofq.log 2>&1
tail -1 $O/c13_t5_phases_nofq.log
( MRB_LIB_VARIANT=_pdls timeout 300 python tools/t5_phase_bench.py $O/c13_t5_phases_pdls.json ) > $O/c13_t5_phases_pdls.log 2>&1
tail -1 $O/c13_t5_phases_pdls.log
( timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c13_bench.json 2> $O/c13_bench.err
( MRB_LIB_VARIANT=_pdls timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c13_bench_pdls.json 2> $O/c13_bench_pdls.err
( timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c13_bench2.json 2> $O/c13_bench2.err
for f in bench bench_pdls bench2; do python -c "
import json; j=json.load(open('$O/c13_$f.json')); print('$f', round(j['ms_per_step'],2), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"; done
( MRB_LIB_VARIANT=_pdls timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -3 ) > $O/c13_pytest_pdls.log 2>&1
tail -2 $O/c13_pytest_pdls.log
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $O/c13_pytest.log 2>&1
tail -4 $O/c13_pytest.log
