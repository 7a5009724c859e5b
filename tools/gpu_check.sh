#!/bin/bash
# One gpurun call: parity tests, smoke, variant sweep, bench lines, ncu launch list + full capture.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== sweep"; timeout 600 python tools/sweep.py --out $OUT/sweep.json 2>&1 | tee $OUT/sweep.txt | tail -70
echo "== sweep split"; timeout 300 python tools/sweep.py --split 1 --filter "float_n12|float_n10" --out $OUT/sweep_split.json 2>&1 | tee $OUT/sweep_split.txt | tail -8
echo "== bench"; for w in cfg2 cfg2s cfg3 cfg4 cfg1; do timeout 600 python bench.py --workload $w --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$w.json; done
echo "== bench full (e2e + cpu)"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench_full.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_cfg2.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_list.log 2>&1; tail -3 $OUT/launches_cfg2.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 3 -c 2 -f -o $OUT/prof_cfg2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1; ls -la $OUT
