#!/bin/bash
# One gpurun call: parity tests, smoke, variant sweep, bench lines of every config, 1-GPU 3D timings.
TAG=${1:-r01b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== sweep"; timeout 600 python tools/sweep.py --out $OUT/sweep.json 2>&1 | tee $OUT/sweep.txt | tail -150
echo "== bench"; for w in cfg2 cfg2s cfg3 cfg4; do timeout 600 python bench.py --workload $w --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$w.json; done
echo "== 3d"; timeout 300 python tools/plan3d_time.py --size 1024 --steps 5 2>&1 | tail -1 | tee $OUT/plan3d_1024.json
timeout 600 python tools/plan3d_time.py --size 2048 --steps 3 2>&1 | tail -1 | tee $OUT/plan3d_2048.json
