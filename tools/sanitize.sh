#!/bin/bash
# compute-sanitizer passes over one eager training step of a 2-layer-deep, full-width model at the QVH shape.
# memcheck sees cudaMalloc granularity only unless the caching allocator is off (second pass, slow).
set -u
O=gpurun_out
mkdir -p $O
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step.py 60 > $O/sanitize_memcheck.log 2>&1
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step.py 8 > $O/sanitize_memcheck_nocache.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_step.py 8 > $O/sanitize_racecheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_step.py 8 > $O/sanitize_synccheck.log 2>&1
grep -H "ERROR SUMMARY" $O/sanitize_*.log
