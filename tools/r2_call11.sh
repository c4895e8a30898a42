#!/bin/bash
# Round 2, GPU call 11: single-lane issue via elect.sync + base-plus-offset descriptors in every tcgen05 kernel (GEMM 1-CTA / 2-CTA,
# attention forward / backward / ViT, LoRA wgrad): correctness, GEMM vs cuBLAS, attention bench, step time, T5 phases in-graph.
set -u
O=gpurun_out
mkdir -p $O
rm -f mr_blip_b200/libmrblip_b200_noepi.so mr_blip_b200/libmrblip_b200_notma.so
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x 2>&1 | tail -30 ) > $O/c11_pytest_kernels.log 2>&1
tail -4 $O/c11_pytest_kernels.log
( timeout 300 python tools/gemm_diag.py $O/c11_diag.json ) > $O/c11_diag.log 2>&1
grep -h "'name'" $O/c11_diag.log | cut -c1-230
( MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py "" tc ) > $O/c11_attn_bench.log 2>&1
grep -v nobias $O/c11_attn_bench.log | cut -c1-120
( timeout 600 python bench.py --steps 10 --warmup 4 --no-eager --no-cpu-baseline ) > $O/c11_bench.json 2> $O/c11_bench.err
cut -c1-300 $O/c11_bench.json; tail -2 $O/c11_bench.err
( timeout 300 python tools/t5_phase_bench.py $O/c11_t5_phases.json ) > $O/c11_t5_phases.log 2>&1
tail -6 $O/c11_t5_phases.log
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $O/c11_pytest.log 2>&1
tail -4 $O/c11_pytest.log
