#!/bin/bash
# r02: 1-GPU validation of the native slab plan + the new bench line
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_full.json 2> $OUT/bench_full.err; tail -c 600 $OUT/bench_full.err; python - <<PY
import json
d = json.loads(open("$OUT/bench_full.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("roofline", d["roofline"])
print("e2e", d["e2e"])
print("cpu", d["cpu_baseline"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: v.get(kk) for kk in ("ms", "gflops", "frac", "hbm_frac_whole_step", "error")}, [(q["axis"], q["frac"]) for q in v.get("per_pass") or []])
print("slab", json.dumps(d.get("slab"))[:1500])
PY
