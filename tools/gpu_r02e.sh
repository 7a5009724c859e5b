#!/bin/bash
# r02: fused two-step variants at the real cfg5 / 1024^3 geometries (flags 13 = discard + evict-first streams + evict-last scratch)
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
V2=( "" "float_n5+6_w16_g32+16_b1_r32x1x1+32x2x1_fused2" "float_n5+6_w16_g32+16_b1_r16x2x1+16x4x1_fused2" "float_n5+6_w16_g16+8_b2_r16x2x1+16x4x1_fused2" )
for V in "${V2[@]}"; do
  B2FFT_PREFER="$V" timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_2048.txt
done
B2FFT_FUSED2=0 timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_2048.txt
V1=( "" "float_n5+5_w16_g32+32_b1_r32x1x1+32x1x1_fused2" "float_n5+5_w16_g32+32_b1_r16x2x1+16x2x1_fused2" "float_n5+5_w16_g16+16_b2_r16x2x1+16x2x1_fused2" )
for V in "${V1[@]}"; do
  B2FFT_PREFER="$V" timeout 300 python tools/axis_time.py --size 1024 --steps 5 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_1024.txt
  B2FFT_PREFER="$V" timeout 300 python tools/axis_time.py --dims 256,1024,1024 --steps 5 --axes 2 2>&1 | cut -c1-200 | tee -a $OUT/axis_cfg3.txt
done
B2FFT_FUSED2=0 timeout 300 python tools/axis_time.py --size 1024 --steps 5 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_1024.txt
B2FFT_FUSED2=0 timeout 300 python tools/axis_time.py --dims 256,1024,1024 --steps 5 --axes 2 2>&1 | cut -c1-200 | tee -a $OUT/axis_cfg3.txt
