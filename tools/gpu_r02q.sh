#!/bin/bash
# r02: N-GPU check of the slab pipeline with the Y pass hidden under the exchange: parity test + 2048^3 timelines
TAG=${1:-r02q}
G=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$G" = "2" ]; then
  echo "== 2-rank pytest"; timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -k two_ranks 2>&1 | tail -5 | tee $OUT/pytest.txt
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29551"
VARS=${3:-"xslabx8z1o0 xslabx8z8o40 xslabx8z8o56 xslabx8z4o40"}
echo "== slab 2048 variants ($G ranks)"
SLAB_TRACE=1 timeout 900 $TR tools/slab_check.py --size 2048 --steps 4 --warmup 2 --exchange $VARS > $OUT/slab2048_g$G.log 2>&1; grep '^{' $OUT/slab2048_g$G.log > $OUT/slab2048_g$G.jsonl; grep -iE "error|Traceback" $OUT/slab2048_g$G.log | sort | uniq -c | head -5
python - <<PY
import json
for l in open("$OUT/slab2048_g$G.jsonl"):
    d = json.loads(l)
    print(d["exchange"], "y%d z%d" % (d["y_chunks"], d["z_chunks"]), "ms=%.2f" % d["ms"], "rt=%.2e" % d.get("roundtrip_rel_l2", -1), "parseval=%.1e" % d["parseval_rel_err"], "delta=%.1e" % d["delta_max_err"], "status", d.get("status"))
    tr = d.get("trace_rank0", [])
    print("   ", " ".join("%s=%.1f" % (k, v) for k, v in tr if not k.startswith("X[") or k.endswith(",7]") or k.endswith(",0]"))[:1800])
PY
if [ "$G" != "2" ]; then
  echo "== parity with the hidden Y pass ($G ranks, 64 x 2048 x 128 gathered)"
  timeout 300 $TR tools/slab_check.py --shape 64,2048,128 --check --steps 0 --exchange xslabx4z4o40 xslabx4z4o0 2>&1 | grep '^{' | tee $OUT/parity_g$G.jsonl | cut -c1-330
fi
