#!/bin/bash
# r02: long-row alias kernel (tmara): per-variant parity + race test + sweep of the long row variants
TAG=${1:-r02ac}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "every_kernel or race_free" 2>&1 | tail -3 | tee $OUT/pytest.txt
echo "== sweep long rows"; timeout 300 python tools/sweep.py --filter "float_n1[34]_w1|double_n1[23]_w1" --mib 2048 --out $OUT/sweep.json 2>&1 | tail -16 | tee $OUT/sweep.txt
echo "== sweep long rows split"; timeout 300 python tools/sweep.py --split 1 --filter "float_n1[34]_w1|double_n1[23]_w1" --mib 2048 --out $OUT/sweep_split.json 2>&1 | tail -16 | tee $OUT/sweep_split.txt
