#!/bin/bash
# Round 2, GPU call 34: last sanity run of the final tree -- smoke(), the kernel tests (incl. the fused residual kernels bit for bit
# on the device), and bench.py exactly as the driver launches it.
set -u
O=gpurun_out
mkdir -p $O
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > $O/c34_smoke.log 2>&1
tail -2 $O/c34_smoke.log
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_splitk_gpu.py -m gpu -q --tb=short 2>&1 | tail -15 ) > $O/c34_pytest.log 2>&1
tail -2 $O/c34_pytest.log
( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 ) > $O/c34_bench.json 2> $O/c34_bench.err
cut -c1-200 $O/c34_bench.json; tail -1 $O/c34_bench.err
