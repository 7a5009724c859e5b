#!/bin/bash
# r02: shared-memory-resident fused two-step kernel: parity + per-axis timings
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_configs_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "cfg5 or cfg3 or every_kernel or 1024 or race" 2>&1 | tail -4 | tee $OUT/pytest.txt
echo "== sweep"; timeout 300 python tools/sweep.py --filter "fused2s" --inner 2048 --mib 2048 --out $OUT/sweep.json 2>&1 | tail -6 | tee $OUT/sweep.txt
echo "== axis 2048 (fused always)"; B2FFT_FUSED2=2 timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_2048.txt
B2FFT_FUSED2=2 B2FFT_PREFER="float_n5+6_w16_g8+8_ks28_fused2s" timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_2048.txt
echo "== axis 1024 (fused always)"; B2FFT_FUSED2=2 timeout 300 python tools/axis_time.py --size 1024 --steps 5 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_1024.txt
B2FFT_FUSED2=2 B2FFT_PREFER="float_n5+5_w16_g8+8_ks32_fused2s" timeout 300 python tools/axis_time.py --size 1024 --steps 5 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_1024.txt
B2FFT_FUSED2=2 timeout 300 python tools/axis_time.py --dims 256,1024,1024 --steps 5 --axes 2 2>&1 | cut -c1-200 | tee -a $OUT/axis_1024.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,launch__registers_per_thread,launch__occupancy_limit_shared_mem
timeout 300 ncu --metrics $M --clock-control none -k regex:fused2s -s 1 -c 1 --csv --log-file $OUT/ncu_z.csv python tools/axis_time.py --dims 2048,64,2048 --axes 4 --steps 2 > $OUT/ncu_z.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("$OUT/ncu_z.csv")) if len(r) > 10]
for r in rows[1:]:
    print("   %-75s %s %s" % (r[-3][:75], r[-1], r[-2]))
PY
