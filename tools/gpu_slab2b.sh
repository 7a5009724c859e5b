#!/bin/bash
TAG=${1:-slab2b}
G=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29543"
echo "== timing 1024^3 on $G ranks"
timeout 300 $RUN tools/slab_check.py --size 1024 --check --steps 5 --warmup 2 --exchange xslabx8c0 xslabx8c1 xslabx8c2 xslabx8c4 xslabx4c2 xslabx16c2 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab1024_g$G.txt | cut -c1-330
echo "== timing 2048^3 on $G ranks"
timeout 400 $RUN tools/slab_check.py --size 2048 --steps 3 --warmup 1 --exchange xslabx8c2 xslabx8c1 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab2048_g$G.txt | cut -c1-330
echo "== bench cfg2/cfg3 1 GPU"; for w in cfg2 cfg3; do timeout 300 python bench.py --workload $w --no-e2e --no-cpu --steps 50 2>&1 | tail -1 | cut -c1-150; done
