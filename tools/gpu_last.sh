#!/bin/bash
TAG=${1:-last}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench"; for w in cfg2 cfg2s cfg3 cfg4; do timeout 600 python bench.py --workload $w --no-e2e --no-cpu 2>&1 | tail -1 > $OUT/bench_$w.json; cut -c1-150 $OUT/bench_$w.json; done
echo "== large 1D"; timeout 300 python tools/time_plan.py 32768:8192 65536:4096 262144:1024 1048576:256 4194304:64 16777216:16 134217728:2 2>&1 | tee $OUT/time_large.txt | cut -c1-150
timeout 300 python tools/time_plan.py 65536:1024 1048576:64 --dtype complex128 2>&1 | tee -a $OUT/time_large.txt | cut -c1-150
