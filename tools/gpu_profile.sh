#!/bin/bash
# ncu evidence for profiles/: launch lists (kernel share of the step) + full captures of the dominant kernels.
TAG=${1:-r01p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for w in cfg2 cfg3 cfg4; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 6 -c 12 --csv --log-file $OUT/launches_$w.csv python bench.py --workload $w --steps 4 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_list_$w.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 3 -c 1 -f -o $OUT/prof_cfg2 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_cfg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 6 -c 2 -f -o $OUT/prof_cfg3 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_cfg3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 9 -c 3 -f -o $OUT/prof_cfg4 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_cfg4.log 2>&1
# keep only CSV exports of the captures (the .ncu-rep files are tens of MB; gpurun_out is capped at 64 MiB)
for w in cfg2 cfg3 cfg4; do
  ncu -i $OUT/prof_$w.ncu-rep --page raw --csv > $OUT/prof_${w}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_$w.ncu-rep --page details --csv > $OUT/prof_${w}_details.csv 2>/dev/null
done
ncu -i $OUT/prof_cfg2.ncu-rep --page source --csv > $OUT/prof_cfg2_source.csv 2>/dev/null
rm -f $OUT/prof_cfg3.ncu-rep $OUT/prof_cfg4.ncu-rep $OUT/prof_cfg2.ncu-rep
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks.csv &
SMI=$!
for w in cfg2 cfg2s cfg3 cfg4 cfg1; do timeout 600 python bench.py --workload $w --no-e2e --no-cpu 2>&1 | tail -1 > $OUT/bench_$w.json; done
timeout 900 python bench.py 2>&1 | tail -1 > $OUT/bench_full.json
kill $SMI
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_reference.json
ls -la $OUT
