#!/bin/bash
# r02 final 1-GPU evidence: parity suite, smoke, examples, full bench line + reference arm, per-config lines,
# published-table shapes, ncu launch list + full capture of the dominant kernels (cfg2 row kernel, cfg5 fused2p passes)
TAG=${1:-r02x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== examples"; timeout 200 python examples/cuda_basic.py 2>&1 | tail -3 | tee $OUT/example_cuda_basic.txt
timeout 200 python examples/multi_gpu_slab.py 2>&1 | tail -3 | tee $OUT/example_slab.txt
echo "== bench (default)"; timeout 900 python bench.py 2> $OUT/bench_full.err | tail -1 > $OUT/bench_full.json; cut -c1-300 $OUT/bench_full.json; tail -2 $OUT/bench_full.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $OUT/bench_reference.json; cut -c1-200 $OUT/bench_reference.json
echo "== bench cfg5 1 GPU"; timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 > $OUT/bench_cfg5_1gpu.json; cut -c1-300 $OUT/bench_cfg5_1gpu.json
echo "== published table"; timeout 600 python tools/published_table.py --out $OUT/published_table.md 2>&1 | tail -3
echo "== ncu launch list (default bench, dominant kernel share)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 6 -c 12 --csv --log-file $OUT/launches_cfg2.csv python bench.py --workload cfg2 --steps 4 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_list_cfg2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fused2p|tile_fft" -c 9 --csv --log-file $OUT/launches_cfg5.csv python tools/axis_time.py --size 2048 --steps 2 --axes 7 > $OUT/ncu_list_cfg5.log 2>&1
tail -12 $OUT/launches_cfg5.csv | cut -c1-400
echo "== ncu full: cfg2 row kernel + fused2p Y pass"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 3 -c 1 -f -o $OUT/prof_cfg2 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_cfg2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2p -s 1 -c 1 -f -o $OUT/prof_fused2p_y python tools/axis_time.py --dims 64,2048,2048 --axes 2 --steps 2 > $OUT/ncu_full_fused2p.log 2>&1
for w in cfg2 fused2p_y; do
  ncu -i $OUT/prof_$w.ncu-rep --page raw --csv > $OUT/prof_${w}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_$w.ncu-rep --page details --csv > $OUT/prof_${w}_details.csv 2>/dev/null
done
rm -f $OUT/prof_cfg2.ncu-rep $OUT/prof_fused2p_y.ncu-rep
ls $OUT
