#!/bin/bash
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== chunk sweep"
for w in cfg3 cfg4; do for mb in 0 8 16 32 48 64 96; do
  B2FFT_L2_CHUNK_MB=$mb timeout 300 python bench.py --workload $w --no-e2e --no-cpu --steps 50 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$w chunk=${mb}MB', d['value'], 'GFLOP/s', d['ms_per_step'], 'ms', 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])" | tee -a $OUT/chunks.txt
done; done
echo "== bench defaults"; for w in cfg2 cfg2s cfg1; do timeout 300 python bench.py --workload $w --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$w.json | cut -c1-400; done
echo "== small-N sweep"; timeout 300 python tools/sweep.py --filter "_n[1-5]_w1" --out $OUT/sweep_small.json 2>&1 | tee $OUT/sweep_small.txt
echo "== ncu cfg3 chunk=32 (dram bytes per kernel)"
B2FFT_L2_CHUNK_MB=32 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 16 --csv --log-file $OUT/ncu_cfg3_chunk32.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_cfg3.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/ncu_cfg3_chunk32.csv")))
hdr = [i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
for r in rows[hdr+1:]:
    print(r[4][:70], r[-3], r[-2], r[-1])
PY
