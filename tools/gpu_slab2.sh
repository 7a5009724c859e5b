#!/bin/bash
TAG=${1:-slab2}
G=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest dist"; timeout 600 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -5 | tee $OUT/pytest_dist.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29543"
echo "== timing 1024^3 on $G ranks"
timeout 300 $RUN tools/slab_check.py --size 1024 --steps 5 --warmup 2 --exchange xslab xslabx4 xslabx16 ncclx8 p2p 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab1024_g$G.txt | cut -c1-330
echo "== timing 2048^3 on $G ranks"
timeout 400 $RUN tools/slab_check.py --size 2048 --steps 3 --warmup 1 --exchange xslab ncclx8 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab2048_g$G.txt | cut -c1-330
