#!/bin/bash
TAG=${1:-iter3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== bench"; for w in cfg2 cfg2s cfg3 cfg4; do timeout 600 python bench.py --workload $w --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$w.json | cut -c1-160; done
echo "== axis 1024"; timeout 300 python tools/axis_time.py --size 1024 2>&1 | tee $OUT/axis_1024.txt
echo "== axis 2048"; timeout 600 python tools/axis_time.py --size 2048 --steps 3 2>&1 | tee $OUT/axis_2048.txt
for v in float_n11_w8_g1_b1_r16x16x8x1 float_n11_w4_g1_b1_r16x16x8x1_tmac1 float_n11_w4_g1_b1_r16x16x8x1_tmac2; do
  echo "-- prefer $v"; B2FFT_PREFER=$v timeout 600 python tools/axis_time.py --size 2048 --steps 3 --axes 2,4 2>&1 | tee -a $OUT/axis_2048.txt
done
echo "== large"; timeout 600 python tools/time_plan.py 8192:32768 16384:16384 65536:4096 1048576:256 4194304:64 2>&1 | tee $OUT/time_large.txt
