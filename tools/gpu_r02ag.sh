#!/bin/bash
# r02: four-elements-per-lane shuffle rows (16-byte loads and stores): per-variant parity + sweeps
TAG=${1:-r02ag}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== parity"; timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "every_kernel" 2>&1 | tail -2 | tee $OUT/pytest.txt
echo "== sweep"; timeout 200 python tools/sweep.py --filter "_shfl|float_n[67]_w1_g" --mib 2048 --iters 5 --out $OUT/sweep.json 2>&1 | tail -14 | tee $OUT/sweep.txt
echo "== sweep split"; timeout 200 python tools/sweep.py --split 1 --filter "_shfl|float_n[67]_w1_g" --mib 2048 --iters 5 --out $OUT/sweep_split.json 2>&1 | tail -14 | tee $OUT/sweep_split.txt
