#!/bin/bash
# r02: short-row shuffle kernel: parity suite + sweep of the row variants N = 4..64 (interleaved and split) + published-table rows
TAG=${1:-r02aa}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_api_contract.py tests/test_configs_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest.txt
echo "== sweep rows"; timeout 300 python tools/sweep.py --filter "float_n[2-6]_w1_" --mib 2048 --out $OUT/sweep.json 2>&1 | tail -12 | tee $OUT/sweep.txt
echo "== sweep rows split"; timeout 300 python tools/sweep.py --split 1 --filter "float_n[2-6]_w1_" --mib 2048 --out $OUT/sweep_split.json 2>&1 | tail -12 | tee $OUT/sweep_split.txt
echo "== published table"; timeout 600 python tools/published_table.py --out $OUT/published_table.md 2>&1 | tail -2; grep -A14 "2 GiB" $OUT/published_table.md | cut -c1-140
