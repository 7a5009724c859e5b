#!/bin/bash
# r02: 2-GPU timeline of the slab exchange pipeline (SLAB_TRACE), grid-cap and chunking variants at 2048^3
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
SLAB_TRACE=1 timeout 900 $TR tools/slab_check.py --size 2048 --steps 3 --exchange xslabx8 xslabx8z4 xslabx8c0 xslabx8c6 xslabx4z4c6 xslabx16z4 2>&1 | grep '^{' | tee $OUT/slab2048.jsonl | cut -c1-200
python - <<PY
import json
for l in open("$OUT/slab2048.jsonl"):
    d = json.loads(l)
    print(d["exchange"], "y%d z%d" % (d["y_chunks"], d["z_chunks"]), "ms=%.2f" % d["ms"], "status", d.get("status"))
    print("   ", " ".join("%s=%.1f" % (k, v) for k, v in d.get("trace_rank0", [])))
PY
