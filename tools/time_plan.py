#!/usr/bin/env python
"""Times Plan.execute for arbitrary shapes (out-of-place unless --inplace) and prints one JSON line each:
   python tools/time_plan.py 1048576:64 4096x4096:4 --dtype complex64 --steps 20
Shape syntax: ZxYxX:batch (numpy order).  Reports GFLOP/s and the per-pass HBM rate (2*bytes per pass)."""
import argparse, json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pyfft_b200.cuda import Plan

ap = argparse.ArgumentParser()
ap.add_argument("cases", nargs="+")
ap.add_argument("--dtype", default="complex64")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--inplace", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
tdt = {"complex64": torch.complex64, "complex128": torch.complex128}[args.dtype]
for case in args.cases:
    shp, _, b = case.partition(":")
    shape = tuple(int(s) for s in shp.split("x"))
    batch = int(b or 1)
    n = int(np.prod(shape))
    a = torch.empty(n * batch, dtype=tdt, device=dev)
    torch.view_as_real(a).normal_()
    out = a if args.inplace else torch.empty_like(a)
    plan = Plan(shape, dtype=np.dtype(args.dtype), stream=torch.cuda.current_stream(), wait_for_finish=False)
    for _ in range(3):
        plan.execute(a, out, batch=batch) if not args.inplace else plan.execute(a, batch=batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.execute(a, out, batch=batch) if not args.inplace else plan.execute(a, batch=batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    npass = len(plan.passes)
    print(json.dumps({"shape": shape, "batch": batch, "dtype": args.dtype, "inplace": args.inplace, "ms": round(ms, 4),
                      "gflops": round(5 * n * math.log2(n) * batch / (ms * 1e-3) / 1e9, 1), "passes": npass,
                      "gbs_per_pass": round(npass * 2 * n * batch * a.element_size() / (ms * 1e-3) / 1e9, 1),
                      "plan": plan.passes}), flush=True)
    del a, out, plan
