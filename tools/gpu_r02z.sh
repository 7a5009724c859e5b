#!/bin/bash
# r02 final N-GPU bench line (cfg2 weak scaling + slab record) under torchrun
TAG=${1:-r02z}; G=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29561"
timeout 700 $TR bench.py --gpus $G --steps 20 --warmup 5 > $OUT/bench_${G}gpu.json 2> $OUT/bench_${G}gpu.err; tail -c 400 $OUT/bench_${G}gpu.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_${G}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"])
print("slab", json.dumps(d.get("slab"))[:2500])
PY
