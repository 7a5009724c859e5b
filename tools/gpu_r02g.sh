#!/bin/bash
# r02: 2-GPU run: two-rank parity test, NVLink store/pull microbenchmark, slab timings (chunking variants), bench --gpus 2
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== two-rank parity"; timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_dist.txt
echo "== nvlink"; timeout 300 ./experiments/nvlink_store_bw 2 2048 > $OUT/nvlink_2048.txt 2>&1; cat $OUT/nvlink_2048.txt | head -50
timeout 300 ./experiments/nvlink_store_bw 2 8192 > $OUT/nvlink_8192.txt 2>&1; grep bulk $OUT/nvlink_8192.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
echo "== slab 1024 / 2048"
timeout 900 $TR tools/slab_check.py --size 1024 --steps 5 --exchange xslabx8 xslabx8z2 xslabx8z4 xslabx8c2 xslabx8c4 xslabx4z2 2>&1 | grep '^{' | tee $OUT/slab1024.jsonl | cut -c1-330
timeout 900 $TR tools/slab_check.py --size 2048 --steps 3 --exchange xslabx8 xslabx8z2 xslabx8z4 xslabx16z2 2>&1 | grep '^{' | tee $OUT/slab2048.jsonl | cut -c1-330
echo "== bench --gpus 2"
timeout 1200 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; tail -c 400 $OUT/bench_2gpu.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_2gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"])
print("slab", json.dumps(d.get("slab"))[:1500])
PY
