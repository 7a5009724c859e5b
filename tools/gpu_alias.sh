#!/bin/bash
TAG=${1:-alias}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest variants"; timeout 400 python -m pytest tests/test_parity_gpu.py -x -q -k "every_kernel_variant or race_free" 2>&1 | tail -4 | tee $OUT/pytest.txt
echo "== sweep"; timeout 200 python tools/sweep.py --filter "float_n11_w4|float_n10_w8_g1_b1_r32x32x1x1_tmac2|tmaca|double_n10_w4" --out $OUT/sweep.json 2>&1 | tee $OUT/sweep.txt | tail -14
echo "== Y pass 2048"; B2FFT_PREFER=float_n11_w4_g1_b2_r16x16x8x1_tmaca timeout 200 python tools/axis_time.py --size 2048 --steps 3 --axes 2 2>&1 | tee $OUT/axis2048_alias.txt | cut -c1-200
echo "== cfg3"; B2FFT_PREFER=float_n10_w8_g1_b2_r16x16x4x1_tmaca timeout 200 python bench.py --workload cfg3 --no-e2e --no-cpu --steps 50 2>&1 | tail -1 | tee $OUT/bench_cfg3_alias.json | cut -c1-150
