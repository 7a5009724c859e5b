#!/bin/bash
# r02: full ncu capture (source-level stall samples) of the fused2p Z pass
TAG=${1:-r02m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
P2=float_n5+6_w16_g16+16_ks28_fused2p,float_n5+5_w16_g16+16_ks32_fused2p
B2FFT_PREFER=$P2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fused2p -s 1 -c 1 -o $OUT/fused2p_z python tools/axis_time.py --dims 2048,64,2048 --axes 4 --steps 2 > $OUT/ncu_z.log 2>&1
tail -3 $OUT/ncu_z.log
ls -la $OUT
