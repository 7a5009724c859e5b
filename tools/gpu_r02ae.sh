#!/bin/bash
# r02: split-layout rows with bulk-copy output planes: parity + cfg2s with and without
TAG=${1:-r02ae}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_api_contract.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.txt
echo "== cfg2s bulk"; timeout 300 python bench.py --workload cfg2s --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" | tee $OUT/cfg2s_bulk.txt
echo "== cfg2s plain stores"; B2FFT_SPLIT_BULK=0 timeout 300 python bench.py --workload cfg2s --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" | tee $OUT/cfg2s_plain.txt
echo "== sweep split rows"; timeout 300 python tools/sweep.py --split 1 --filter "_tma[12]$" --mib 2048 --out $OUT/sweep_split.json 2>&1 | tail -8 | tee $OUT/sweep_split.txt
