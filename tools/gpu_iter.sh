#!/bin/bash
# One tuning iteration on the GPU: full parity suite, variant sweep, bench lines, large/3D timings.
TAG=${1:-iter}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== sweep"; timeout 600 python tools/sweep.py --filter "${SWEEP_FILTER:-.*}" --out $OUT/sweep.json 2>&1 | tee $OUT/sweep.txt | tail -150
echo "== bench"; for w in cfg2 cfg2s cfg3 cfg4; do timeout 600 python bench.py --workload $w --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$w.json; done
echo "== large"; timeout 600 python tools/time_plan.py 65536:4096 1048576:256 16777216:16 4096x4096:16 1024x1024x1024:1 2>&1 | tee $OUT/time_large.txt
timeout 600 python tools/plan3d_time.py --size 2048 --steps 3 2>&1 | tail -1 | tee $OUT/plan3d_2048.json
