#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -k "blip2_t5" 2>&1 | tail -25 ) > $O/pytest.log 2>&1
tail -12 $O/pytest.log
( timeout 300 python tools/determinism_check.py ) > $O/determinism.log 2>&1; tail -12 $O/determinism.log
