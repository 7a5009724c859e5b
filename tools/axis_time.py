#!/usr/bin/env python
"""Times each axis pass of a 3D transform separately (in place), through b2fft_plan_create_ex's axes mask:
   python tools/axis_time.py --size 2048 [--dtype complex64] [--steps 3]
B2FFT_PREFER=<variant,...> selects alternative kernels.  Prints one JSON line per axis."""
import argparse, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pyfft_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--dims", default="")          # "z,y,x" overrides --size
ap.add_argument("--dtype", default="complex64")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--axes", default="1,2,4")
args = ap.parse_args()
lib = _lib.load()
dims = [int(v) for v in args.dims.split(",")] if args.dims else [args.size] * 3
nz, ny, nx = dims
prec = _lib.F32 if args.dtype == "complex64" else _lib.F64
tdt = torch.complex64 if prec == _lib.F32 else torch.complex128
a = torch.empty(nz * ny * nx, dtype=tdt, device="cuda:0")
ar = torch.view_as_real(a)
step = max(1, a.numel() // 64)
for i in range(0, a.numel(), step):
    ar[i:i + step].normal_()
stream = torch.cuda.current_stream().cuda_stream
for mask in [int(m) for m in args.axes.split(",")]:
    h = ctypes.c_void_p()
    _lib.check(lib.b2fft_plan_create_ex(ctypes.byref(h), (ctypes.c_int64 * 3)(nx, ny, nz), mask, prec, _lib.INTERLEAVED,
                                        1, 1.0, 1, 0, 0.0, 0))
    buf = ctypes.create_string_buffer(4096)
    lib.b2fft_plan_describe(h, buf, len(buf))
    need = ctypes.c_size_t(0)
    lib.b2fft_plan_workspace_bytes(h, 1, ctypes.byref(need))
    ws = torch.empty(need.value, dtype=torch.uint8, device="cuda:0") if need.value else None
    if ws is not None:
        lib.b2fft_plan_set_workspace(h, ws.data_ptr(), need.value)
    run = lambda: _lib.check(lib.b2fft_execute(h, a.data_ptr(), None, a.data_ptr(), None, 0, 1, stream))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    npass = len(buf.value.decode().strip().splitlines())
    print(json.dumps({"dims": dims, "mask": mask, "ms": round(ms, 3), "gbs_per_pass": round(npass * 2 * a.numel() * a.element_size() / ms / 1e6, 1),
                      "plan": buf.value.decode().strip().splitlines()}), flush=True)
    lib.b2fft_plan_destroy(h)
    a.mul_(1e-3)     # keep magnitudes finite across repeated unnormalised transforms
