#!/bin/bash
# r02: streamed in-place fused two-step kernel (fused2p): parity + sweep + per-axis timings + ncu
TAG=${1:-r02l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "every_kernel" 2>&1 | tail -4 | tee $OUT/pytest.txt
echo "== sweep 16KiB pitch"; timeout 300 python tools/sweep.py --filter "fused2p|n5\+6_w16_g16\+16_ks28_fused2s|n11_w4" --inner 2048 --mib 2048 --out $OUT/sweep.json 2>&1 | tail -8 | tee $OUT/sweep.txt
P2=float_n5+6_w16_g16+16_ks28_fused2p,float_n5+5_w16_g16+16_ks32_fused2p
echo "== axis 2048 (fused2p always)"; B2FFT_PREFER=$P2 B2FFT_FUSED2=2 timeout 300 python tools/axis_time.py --size 2048 --steps 3 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_2048.txt
echo "== axis 1024 (fused2p always)"; B2FFT_PREFER=$P2 B2FFT_FUSED2=2 timeout 300 python tools/axis_time.py --size 1024 --steps 5 --axes 2,4 2>&1 | cut -c1-200 | tee -a $OUT/axis_1024.txt
B2FFT_PREFER=$P2 B2FFT_FUSED2=2 timeout 300 python tools/axis_time.py --dims 256,1024,1024 --steps 5 --axes 2 2>&1 | cut -c1-200 | tee -a $OUT/axis_1024.txt
B2FFT_PREFER=$P2 B2FFT_FUSED2=2 timeout 300 python tools/axis_time.py --dims 256,2048,2048 --steps 5 --axes 2 2>&1 | cut -c1-200 | tee -a $OUT/axis_slabY.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,launch__registers_per_thread,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none -k regex:fused2p -s 1 -c 1 --csv --log-file $OUT/ncu_z.csv python tools/axis_time.py --dims 2048,64,2048 --axes 4 --steps 2 > $OUT/ncu_z.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("$OUT/ncu_z.csv")) if len(r) > 10]
for r in rows[1:]:
    print("   %-75s %s %s" % (r[-3][:75], r[-1], r[-2]))
PY
