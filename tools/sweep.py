#!/usr/bin/env python
"""Times every compiled kernel variant (one pass each) and prints achieved GB/s.

    python tools/sweep.py [--filter REGEX] [--mib 1024] [--out gpurun_out/sweep.json]

Algorithmic bytes = read + write of every element once.  Row variants (W=1) run on [tiles][N];
column variants on [outer][N][inner] with inner = --inner (default 1024).  Tuning tool only.
"""
import argparse
import ctypes
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filter", default=".*")
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--inner", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--split", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    args = ap.parse_args()
    import torch
    from pyfft_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    peak = 6543.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    nbytes = args.mib << 20
    a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    a32 = a.view(torch.float32)
    a32.normal_()
    # reference copy bandwidth with the same buffers (what MEASURED_PEAKS measures)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        b.copy_(a)
    ev0.record()
    for _ in range(10):
        b.copy_(a)
    ev1.record()
    torch.cuda.synchronize()
    copy_gbs = 2 * nbytes * 10 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    print("torch copy_: %.0f GB/s (MEASURED_PEAKS %.0f)" % (copy_gbs, peak), flush=True)
    results = {"copy_gbs": copy_gbs, "variants": []}
    buf = ctypes.create_string_buffer(256)
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(lib.b2fft_num_variants()):
        lib.b2fft_variant_info(i, buf, len(buf))
        f = buf.value.decode().split()
        name, prec, lg, W, G, E, S, threads, smem, minb, occ = f[0], *map(int, f[1:])
        if not re.search(args.filter, name):
            continue
        n = 1 << lg
        csize = 8 if prec == 0 else 16
        inner = 1 if W == 1 else max(W, args.inner)
        if W > 1 and n * inner * csize > nbytes:
            continue
        n_el = nbytes // csize
        lines = n_el // n                     # number of columns
        if W > 1:
            lines = (lines // inner) * inner
        if lines == 0:
            continue
        n_tiles = lines // W
        n_used = lines * n
        if args.split:
            half = nbytes // 2
            in0, in1, out0, out1 = a.data_ptr(), a.data_ptr() + half, b.data_ptr(), b.data_ptr() + half
            n_tiles //= 2
            n_used //= 2
            if W > 1:
                n_tiles = (n_tiles // (inner // W)) * (inner // W)
                n_used = n_tiles * W * n
        else:
            in0, in1, out0, out1 = a.data_ptr(), None, b.data_ptr(), None

        def run():
            _lib.check(lib.b2fft_run_variant(i, in0, in1, out0, out1, args.split, 0, n_tiles, inner, 0, stream))
        try:
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(args.iters):
                run()
            ev1.record()
            torch.cuda.synchronize()
        except Exception as exc:
            print("%-44s FAILED %s" % (name, exc), flush=True)
            continue
        ms = ev0.elapsed_time(ev1) / args.iters
        gbs = 2 * csize * n_used / (ms * 1e-3) / 1e9
        print("%-44s thr=%4d smem=%6d occ=%2d  %8.3f ms  %7.0f GB/s  %5.1f%% of peak" % (
            name, threads, smem, occ, ms, gbs, 100 * gbs / peak), flush=True)
        results["variants"].append({"name": name, "threads": threads, "smem": smem, "occ": occ, "ms": ms,
                                    "gbs": gbs, "frac": gbs / peak, "inner": inner, "split": args.split})
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
