#!/bin/bash
# Round evidence in one gpurun call: parity suite, smoke, bench lines (+ e2e, CPU baseline, reference arm),
# ncu launch lists and full captures of the dominant kernels, clocks during the bench runs.
TAG=${1:-r01p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks.csv &
SMI=$!
echo "== bench"; for w in cfg2 cfg2s cfg3 cfg4 cfg1; do timeout 600 python bench.py --workload $w --no-e2e --no-cpu 2>&1 | tail -1 > $OUT/bench_$w.json; cut -c1-140 $OUT/bench_$w.json; done
timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 3 2>&1 | tail -1 > $OUT/bench_cfg5_1gpu.json; cut -c1-140 $OUT/bench_cfg5_1gpu.json
timeout 900 python bench.py 2>&1 | tail -1 > $OUT/bench_full.json; cut -c1-200 $OUT/bench_full.json
kill $SMI
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_reference.json; cut -c1-200 $OUT/bench_reference.json
echo "== timings"; timeout 300 python tools/time_plan.py 8192:32768 16384:16384 65536:4096 1048576:256 4194304:64 16777216:16 2>&1 | tee $OUT/time_large.txt | cut -c1-200
timeout 300 python tools/axis_time.py --size 1024 2>&1 | tee $OUT/axis_1024.txt | cut -c1-200
timeout 120 python tools/sweep.py --split 1 --filter "float_n12|float_n10_w1|float_n11_w1" --out $OUT/sweep_split.json 2>&1 | tee $OUT/sweep_split.txt | tail -12
echo "== ncu launch lists"
for w in cfg2 cfg3 cfg4; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 6 -c 12 --csv --log-file $OUT/launches_$w.csv python bench.py --workload $w --steps 4 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_list_$w.log 2>&1
done
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 3 -c 1 -f -o $OUT/prof_cfg2 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_cfg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 6 -c 2 -f -o $OUT/prof_cfg3 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_cfg3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 9 -c 3 -f -o $OUT/prof_cfg4 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_cfg4.log 2>&1
for w in cfg2 cfg3 cfg4; do
  ncu -i $OUT/prof_$w.ncu-rep --page raw --csv > $OUT/prof_${w}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_$w.ncu-rep --page details --csv > $OUT/prof_${w}_details.csv 2>/dev/null
done
ncu -i $OUT/prof_cfg2.ncu-rep --page source --csv > $OUT/prof_cfg2_source.csv 2>/dev/null
rm -f $OUT/prof_cfg3.ncu-rep $OUT/prof_cfg4.ncu-rep $OUT/prof_cfg2.ncu-rep
ls -la $OUT | head -50
