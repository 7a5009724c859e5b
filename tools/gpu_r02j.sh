#!/bin/bash
# r02: 8-GPU run: slab exchange pipeline variants with timelines, host<->device copy scaling, full bench line
TAG=${1:-r02j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
G=${2:-8}
nvidia-smi topo -m > $OUT/topo.txt 2>&1
lscpu | head -25 > $OUT/lscpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29551"
echo "== slab 2048 variants"
SLAB_TRACE=1 timeout 900 $TR tools/slab_check.py --size 2048 --steps 4 --warmup 2 --exchange xslabx8 xslabx8z2 xslabx8z4 xslabx8z4c2 xslabx8z4c4 xslabx8z4c6 xslabx16z4 xslabx4z2 xslabx8z8 2>&1 | grep '^{' > $OUT/slab2048.jsonl
python - <<PY
import json
for l in open("$OUT/slab2048.jsonl"):
    d = json.loads(l)
    print(d["exchange"], "y%d z%d" % (d["y_chunks"], d["z_chunks"]), "ms=%.2f" % d["ms"], "rt=%.2e" % d.get("roundtrip_rel_l2", -1), "status", d.get("status"))
    print("   ", " ".join("%s=%.1f" % (k, v) for k, v in d.get("trace_rank0", []))[:1500])
PY
echo "== parity 512 (8 ranks, gathered)"
timeout 600 $TR tools/slab_check.py --size 512 --check --steps 0 --exchange xslabx8z2 2>&1 | grep '^{' | tee $OUT/slab512_parity.jsonl | cut -c1-400
echo "== pcie scaling"
timeout 300 $TR tools/pcie_scaling.py > $OUT/pcie_8.json 2>&1; tail -1 $OUT/pcie_8.json | cut -c1-600
timeout 300 $TR tools/pcie_scaling.py --bind > $OUT/pcie_8_bind.json 2>&1; tail -1 $OUT/pcie_8_bind.json | cut -c1-300
timeout 300 python tools/pcie_scaling.py > $OUT/pcie_1.json 2>&1; tail -1 $OUT/pcie_1.json | cut -c1-300
echo "== bench --gpus $G"
timeout 1200 $TR bench.py --gpus $G --steps 20 --warmup 5 > $OUT/bench_${G}gpu.json 2> $OUT/bench_${G}gpu.err; tail -c 300 $OUT/bench_${G}gpu.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_${G}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"])
print("slab", json.dumps(d.get("slab"))[:1800])
PY
