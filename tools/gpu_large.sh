#!/bin/bash
TAG=${1:-r01c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest large"; timeout 900 python -m pytest tests/test_large_gpu.py -x -q 2>&1 | tail -15 | tee $OUT/pytest_large.txt
echo "== timing"; timeout 600 python tools/time_plan.py 32768:8192 65536:4096 1048576:256 4194304:64 16777216:16 134217728:2 4096x4096:16 2>&1 | tee $OUT/time_large.txt
timeout 600 python tools/time_plan.py 65536:1024 1048576:64 --dtype complex128 2>&1 | tee -a $OUT/time_large.txt
timeout 600 python tools/time_plan.py 1048576:256 --inplace 2>&1 | tee -a $OUT/time_large.txt
