"""One large 3D transform slab-decomposed over the GPUs of a node (no reference counterpart):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/multi_gpu_slab.py

Rank g fills its z-slab [Z/G][Y][X]; forward() leaves F[kz][ky][kx in block g] on rank g as [Y][Z][X/G]."""
import os
import sys

import numpy
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a source checkout
from pyfft_b200.dist import SlabPlan

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))

n = 512
plan = SlabPlan((n, n, n), dtype=numpy.complex64, exchange="xslab")
plan.slab.zero_()
if rank == 0:
    plan.slab[0, 0, 0] = 1.0                      # a delta at the origin ...
out = plan.forward()                              # ... transforms to all ones
torch.cuda.synchronize()
print("rank %d: x-slab %s, max |F - 1| = %.1e" % (rank, tuple(out.shape), float((out - 1).abs().max())))
back = plan.inverse()                             # x-slabs -> z-slabs
torch.cuda.synchronize()
print("rank %d: round trip error %.1e" % (rank, float((back[0, 0, 0] - (1.0 if rank == 0 else 0.0)).abs())))
plan.close()
if world > 1:
    dist.destroy_process_group()
