#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
