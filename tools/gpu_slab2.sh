#!/bin/bash
TAG=${1:-slab2}
G=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest dist + abi"; timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_api_contract.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_dist.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29543"
echo "== timing 1024^3 on $G ranks (bulk stores)"
timeout 200 $RUN tools/slab_check.py --size 1024 --steps 5 --warmup 2 --exchange xslabx8c0 xslabx8c1 xslabx8c2 xslabx8c3 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab1024_g${G}_bulk.txt | cut -c1-330
echo "== timing 1024^3 on $G ranks (LSU stores)"
B2FFT_BLK_BULK=0 timeout 200 $RUN tools/slab_check.py --size 1024 --steps 5 --warmup 2 --exchange xslabx8c0 xslabx8c2 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab1024_g${G}_lsu.txt | cut -c1-330
echo "== timing 2048^3 on $G ranks"
timeout 300 $RUN tools/slab_check.py --size 2048 --steps 3 --warmup 1 --exchange xslabx8c1 xslabx8c2 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab2048_g${G}_bulk.txt | cut -c1-330
echo "== split"; timeout 300 python bench.py --workload cfg2s --no-e2e --no-cpu --steps 50 2>&1 | tail -1 | cut -c1-700
