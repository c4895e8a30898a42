#!/bin/bash
( timeout 55 python -m pytest tests/test_full_size_gpu.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pf_final.log 2>&1; cat gpurun_out/pf_final.log | cut -c1-250
