#!/bin/bash
# r02: last validation of the tree as committed: full GPU suite, smoke, default bench line, published table
TAG=${1:-r02ad}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench (default)"; timeout 900 python bench.py 2> $OUT/bench_full.err | tail -1 > $OUT/bench_full.json; cut -c1-200 $OUT/bench_full.json; tail -2 $OUT/bench_full.err
echo "== published table"; timeout 600 python tools/published_table.py --out $OUT/published_table.md 2>&1 | tail -1
