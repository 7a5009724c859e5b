#!/bin/bash
TAG=${1:-iter2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== sweep"; timeout 900 python tools/sweep.py --out $OUT/sweep.json 2>&1 | tee $OUT/sweep.txt | tail -3
echo "== bench cfg2 candidates"
for v in "" float_n12_w1_g1_b4_r16x16x16x1 float_n12_w1_g1_b4_r16x16x16x1_tw0 float_n12_w1_g2_b2_r64x64x1x1 float_n12_w1_g4_b1_r64x64x1x1 float_n12_w1_g1_b4_r32x32x4x1 float_n12_w1_g1_b5_r16x16x16x1; do
  B2FFT_PREFER=$v timeout 300 python bench.py --workload cfg2 --no-e2e --no-cpu --steps 50 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['roofline']['frac'], d['config']['plan'])" | tee -a $OUT/cfg2_candidates.txt
done
echo "== bench cfg3 candidates"
for v in "" float_n10_w1_g4_b4_r32x32x1x1 "float_n10_w1_g4_b4_r32x32x1x1,float_n10_w8_g1_b3_r32x32x1x1" "float_n10_w1_g4_b4_r32x32x1x1,float_n10_w4_g1_b4_r32x32x1x1"; do
  B2FFT_PREFER=$v timeout 300 python bench.py --workload cfg3 --no-e2e --no-cpu --steps 50 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['roofline']['frac'], d['config']['plan'])" | tee -a $OUT/cfg3_candidates.txt
done
echo "== ncu"
B2FFT_PREFER=float_n12_w1_g1_b4_r16x16x16x1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 3 -c 1 -f -o $OUT/prof_cfg2_plain python bench.py --workload cfg2 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_plain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 3 -c 1 -f -o $OUT/prof_cfg2_tma python bench.py --workload cfg2 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_tma.log 2>&1
for w in plain tma; do
  ncu -i $OUT/prof_cfg2_$w.ncu-rep --page raw --csv > $OUT/prof_cfg2_${w}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_cfg2_$w.ncu-rep --page source --csv > $OUT/prof_cfg2_${w}_source.csv 2>/dev/null
  rm -f $OUT/prof_cfg2_$w.ncu-rep
done
ls -la $OUT
