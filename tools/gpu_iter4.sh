#!/bin/bash
TAG=${1:-iter4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== bench (features on)"; for w in cfg2 cfg2s cfg3 cfg4; do timeout 600 python bench.py --workload $w --no-e2e --no-cpu --steps 50 2>&1 | tail -1 | tee $OUT/bench_$w.json | cut -c1-150; done
echo "== cfg3 without TMA stores"; B2FFT_TMA_STORE=0 timeout 600 python bench.py --workload cfg3 --no-e2e --no-cpu --steps 50 2>&1 | tail -1 | tee $OUT/bench_cfg3_nots.json | cut -c1-150
echo "== large 1D (bulk on)"; timeout 300 python tools/time_plan.py 65536:4096 262144:1024 1048576:256 4194304:64 16777216:16 2>&1 | tee $OUT/time_large_on.txt | cut -c1-150
echo "== large 1D (bulk off, tma store off)"; B2FFT_FS_BULK=0 B2FFT_TMA_STORE=0 timeout 300 python tools/time_plan.py 65536:4096 262144:1024 1048576:256 4194304:64 16777216:16 2>&1 | tee $OUT/time_large_off.txt | cut -c1-150
echo "== axis 2048 Y (on/off)"; timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 2 2>&1 | tee $OUT/axis2048_on.txt | cut -c1-200
B2FFT_TMA_STORE=0 timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 2 2>&1 | tee $OUT/axis2048_off.txt | cut -c1-200
echo "== axis 1024"; timeout 300 python tools/axis_time.py --size 1024 2>&1 | tee $OUT/axis_1024.txt | cut -c1-200
echo "== c128 2^20"; timeout 300 python tools/time_plan.py 1048576:64 --dtype complex128 2>&1 | tee $OUT/time_large_c128.txt | cut -c1-150
