#!/bin/bash
# ncu full captures of the strided-axis kernels that bound cfg5 (Y: TMA-staged N=2048 x W=4, Z: plain N=2048 x W=8 at 32 MiB pitch)
TAG=${1:-ncu_axes}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 1 -c 1 -f -o $OUT/prof_y2048 python tools/axis_time.py --size 2048 --steps 1 --axes 2 > $OUT/ncu_y.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 1 -c 1 -f -o $OUT/prof_z2048 python tools/axis_time.py --size 2048 --steps 1 --axes 4 > $OUT/ncu_z.log 2>&1
for w in y2048 z2048; do
  ncu -i $OUT/prof_$w.ncu-rep --page raw --csv > $OUT/prof_${w}_raw.csv 2>/dev/null
  rm -f $OUT/prof_$w.ncu-rep
done
tail -3 $OUT/ncu_y.log $OUT/ncu_z.log
ls -la $OUT
