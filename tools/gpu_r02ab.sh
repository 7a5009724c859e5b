#!/bin/bash
TAG=${1:-r02ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.txt
echo "== sweep rows"; timeout 300 python tools/sweep.py --filter "float_n[2-6]_w1_" --mib 2048 --out $OUT/sweep.json 2>&1 | tail -12 | tee $OUT/sweep.txt
echo "== sweep rows split"; timeout 300 python tools/sweep.py --split 1 --filter "float_n[2-6]_w1_shfl|float_n5_w1_g32|float_n4_w1_g32" --mib 2048 --out $OUT/sweep_split.json 2>&1 | tail -8 | tee $OUT/sweep_split.txt
