#!/bin/bash
# 8-GPU x-slab 3D FFT: parity at 512^3 on 8 ranks, timing at 2048^3 on 8 and 4 ranks.
TAG=${1:-slab8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
RUN8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
RUN4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542"
echo "== parity 512^3 on 8 ranks"
timeout 200 $RUN8 tools/slab_check.py --size 512 --check --steps 3 --warmup 1 --exchange xslab xslabx4c4 --out $OUT/slab512_g8.json 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab512_g8.txt | cut -c1-300
echo "== timing 2048^3 on 8 ranks"
timeout 300 $RUN8 tools/slab_check.py --size 2048 --steps 4 --warmup 2 --exchange xslabx8c2 xslabx8c3 xslabx8c4 xslabx16c3 xslabx4c3 ncclx8 --out $OUT/slab2048_g8.json 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab2048_g8.txt | cut -c1-300
echo "== timing 2048^3 on 4 ranks"
timeout 200 $RUN4 tools/slab_check.py --size 2048 --steps 3 --warmup 1 --exchange xslabx8c2 xslabx8c4 --out $OUT/slab2048_g4.json 2>&1 | grep -E "^\{|Error|error" | tee $OUT/slab2048_g4.txt | cut -c1-300
