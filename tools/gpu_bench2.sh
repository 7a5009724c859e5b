#!/bin/bash
TAG=${1:-bench2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
echo "== bench.py --gpus 2 (default workload, as the driver launches it)"
timeout 300 $RUN bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | grep -E "^\{|Error|error" | tee $OUT/bench_cfg2_g2.json | cut -c1-400
echo "== bench.py --gpus 2 --impl reference"
timeout 200 $RUN bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | grep -E "^\{|Error|error" | tee $OUT/bench_ref_g2.json | cut -c1-200
echo "== bench.py --workload cfg5s --gpus 2"
timeout 200 $RUN bench.py --gpus 2 --workload cfg5s --steps 10 --warmup 3 2>&1 | grep -E "^\{|Error|error" | tee $OUT/bench_cfg5s_g2.json | cut -c1-600
