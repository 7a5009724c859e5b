#!/bin/bash
# 8-GPU slab-decomposed 3D FFT: parity at 512^3, timing at 2048^3 for every exchange mode.
TAG=${1:-slab8}
G=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/gpus.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541"
echo "== parity 512^3 on $G ranks"
timeout 240 $RUN tools/slab_check.py --size 512 --check --steps 3 --warmup 1 --exchange p2p nccl p2p-yzx ncclx4 --out $OUT/slab512_g$G.json 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab512_g$G.txt
echo "== timing 2048^3 on $G ranks"
timeout 420 $RUN tools/slab_check.py --size 2048 --steps 3 --warmup 1 --exchange ${EXCH:-p2p p2p-yzx nccl ncclx4 ncclx8} --out $OUT/slab2048_g$G.json 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab2048_g$G.txt
