#!/bin/bash
# r02: first run of the fused two-step strided kernels: parity suite, variant sweep, per-axis timings.
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== sweep fused"; timeout 300 python tools/sweep.py --filter "fused|float_n1[01]_w" --inner 1024 --out $OUT/sweep_fused_1024.json 2>&1 | tee $OUT/sweep_fused_1024.txt | tail -30
timeout 300 python tools/sweep.py --filter "fused|float_n1[01]_w" --inner 4096 --mib 4096 --out $OUT/sweep_fused_4096.json 2>&1 | tee $OUT/sweep_fused_4096.txt | tail -30
echo "== axis 2048"; timeout 300 python tools/axis_time.py --size 2048 --steps 3 2>&1 | tee $OUT/axis_2048.txt | cut -c1-250
B2FFT_PREFER="float_n5+6_w16_g32+16_b1_r32x1x1+32x2x1_fused2" timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 2,4 2>&1 | tee $OUT/axis_2048_t512.txt | cut -c1-250
B2FFT_PREFER="float_n5+6_w16_g16+8_b2_r16x2x1+16x4x1_fused2" timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 2,4 2>&1 | tee $OUT/axis_2048_e16.txt | cut -c1-250
echo "== axis 1024"; timeout 300 python tools/axis_time.py --size 1024 --steps 5 2>&1 | tee $OUT/axis_1024.txt | cut -c1-250
echo "== bench"; for w in cfg3 cfg5; do timeout 600 python bench.py --workload $w --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | tail -1 > $OUT/bench_$w.json; cut -c1-200 $OUT/bench_$w.json; done
