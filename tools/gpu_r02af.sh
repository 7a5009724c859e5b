#!/bin/bash
TAG=${1:-r02af}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== configs + large + dist"; timeout 900 python -m pytest tests/test_configs_gpu.py tests/test_large_gpu.py tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.txt
echo "== bench (default)"; timeout 900 python bench.py 2> $OUT/bench_full.err | tail -1 > $OUT/bench_full.json; cut -c1-160 $OUT/bench_full.json
