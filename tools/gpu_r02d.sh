#!/bin/bash
# r02: fused two-step kernel -- effect of the L2 policy / discard / prefetch flags, DRAM traffic under ncu
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for F in 0 1 2 4 8 9 12 13 15; do
  echo "== flags $F"
  B2FFT_FUSED_FLAGS=$F timeout 120 python tools/axis_time.py --dims 256,2048,2048 --axes 2 --steps 5 2>&1 | cut -c1-120 | tee -a $OUT/flags_y.txt
  B2FFT_FUSED_FLAGS=$F timeout 120 python tools/axis_time.py --dims 2048,64,2048 --axes 4 --steps 5 2>&1 | cut -c1-120 | tee -a $OUT/flags_z.txt
done
for V in "float_n5+6_w16_g32+16_b1_r32x1x1+32x2x1_fused2"; do
  B2FFT_PREFER=$V timeout 120 python tools/axis_time.py --dims 256,2048,2048 --axes 2 --steps 5 2>&1 | cut -c1-200 | tee -a $OUT/flags_y.txt
  B2FFT_PREFER=$V timeout 120 python tools/axis_time.py --dims 2048,64,2048 --axes 4 --steps 5 2>&1 | cut -c1-200 | tee -a $OUT/flags_z.txt
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
for F in 0 15; do
  B2FFT_FUSED_FLAGS=$F timeout 300 ncu --metrics $M --clock-control none -k regex:fused2 -s 1 -c 1 --csv --log-file $OUT/ncu_y_f$F.csv python tools/axis_time.py --dims 256,2048,2048 --axes 2 --steps 2 > $OUT/ncu_y_f$F.log 2>&1
  B2FFT_FUSED_FLAGS=$F timeout 300 ncu --metrics $M --clock-control none -k regex:fused2 -s 1 -c 1 --csv --log-file $OUT/ncu_z_f$F.csv python tools/axis_time.py --dims 2048,64,2048 --axes 4 --steps 2 > $OUT/ncu_z_f$F.log 2>&1
done
python - <<PY
import csv, glob, os
for f in sorted(glob.glob(os.path.join("$OUT", "ncu_*.csv"))):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    print(f)
    for r in rows[1:]:
        print("   %-75s %s %s" % (r[-3][:75], r[-1], r[-2]))
PY
