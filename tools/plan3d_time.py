#!/usr/bin/env python
"""Times a single-GPU in-place 3D Plan (the 1-GPU point of the slab scaling series).
   python tools/plan3d_time.py --size 2048 --steps 3"""
import argparse, json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pyfft_b200.cuda import Plan

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dtype", default="complex64")
args = ap.parse_args()
n = args.size
dev = torch.device("cuda:0")
tdt = torch.complex64 if args.dtype == "complex64" else torch.complex128
fl = torch.float32 if args.dtype == "complex64" else torch.float64
a = torch.empty(n, n, n, dtype=tdt, device=dev)
ar = torch.view_as_real(a)
for z in range(0, n, 64):            # fill in slices (no 64 GiB temporaries)
    ar[z:z + 64].normal_()
plan = Plan((n, n, n), dtype=np.dtype(args.dtype), stream=torch.cuda.current_stream(), wait_for_finish=False)
e_in = float((torch.view_as_real(a[:8]).double() ** 2).sum().item())
probe = a[:, 5, 7].clone()           # one z-line before the transform is not enough to check; use delta + Parseval below
plan.execute(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    plan.execute(a)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
N = float(n) ** 3
rec = {"n": n, "world": 1, "impl": "Plan in-place", "dtype": args.dtype, "ms": ms, "gflops": 5 * N * math.log2(N) / (ms * 1e-3) / 1e9,
       "plan": plan.passes, "hbm_gbs_3pass": 3 * 2 * N * a.element_size() / (ms * 1e-3) / 1e9}
# closed form: delta -> ones; plane wave -> delta
a.zero_()
a[0, 0, 0] = 1
plan.execute(a)
torch.cuda.synchronize()
rec["delta_max_err"] = max(float((a[z0:z0 + 64] - 1).abs().max().item()) for z0 in range(0, n, 64))
plan.execute(a, inverse=True)
torch.cuda.synchronize()
rec["delta_roundtrip_err"] = float((a[0, 0, 0] - 1).abs().item()) + max(float(a[z0:z0 + 64].abs().sum().item()) for z0 in range(64, n, 64))
print(json.dumps(rec))
