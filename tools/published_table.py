#!/usr/bin/env python
"""The reference's own performance table (doc/source/index.rst:319-374) re-measured on this GPU with its own protocol
(test/test_performance.py:3-32): shapes of test_performance.py:40-44, a 32 MB buffer per shape (batch = 32 MiB /
(x*y*z*itemsize)), out-of-place executes, one warm-up then 10 timed iterations, GFLOPS = 5e-9 * log2(x*y*z) * x*y*z *
batch / t.  A 32 MB working set fits the 126 MB L2 of a B200, so these rows measure L2 + launch latency, not HBM; the
second block therefore repeats every shape with a 2 GiB buffer (larger than L2), which is what bench.py's configs use.

    python tools/published_table.py [--out profiles/r02_published_table.md]
"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SHAPES = [(16,), (1024,), (8192,), (16, 16), (128, 128), (1024, 1024), (8, 8, 64), (16, 16, 16), (16, 16, 128),
          (32, 32, 128), (128, 128, 128)]
# pyfft sp / dp GFLOPS published for the Tesla C2050, keyed by the doc's [x, y, z] label (index.rst:352-374)
PUBLISHED = {(16,): (91.4, 37.8), (1024,): (254.0, 28.4), (8192,): (117.3, 29.3), (16, 16): (106.7, 43.4),
             (128, 128): (187.2, 47.4), (1024, 1024): (168.5, 27.7), (16, 16, 16): (117.6, 47.0),
             (32, 32, 128): (163.9, 58.8), (128, 128, 128): (184.7, 44.6)}


def measure(shape, dtype, buffer_mib, iterations=10):
    from pyfft_b200.cuda import Plan
    size = int(np.prod(shape))
    itemsize = np.dtype(dtype).itemsize
    batch = (buffer_mib << 20) // (size * itemsize)
    if batch == 0:
        return None
    tdt = torch.complex64 if dtype == np.complex64 else torch.complex128
    fl = torch.float32 if dtype == np.complex64 else torch.float64
    a = torch.view_as_complex(torch.randn(size * batch, 2, dtype=fl, device="cuda:0"))
    b = torch.empty_like(a)
    plan = Plan(shape, dtype=dtype, wait_for_finish=False, stream=torch.cuda.current_stream())
    plan.execute(a, b, batch=batch)                       # warming up (test_performance.py:26)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iterations):
        plan.execute(a, b, batch=batch)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / iterations
    gflop = 5.0e-9 * math.log2(size) * size * batch
    return batch, t * 1e3, gflop / t, len(plan.passes) * 2.0 * itemsize * size * batch / t / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_published_table.md"))
    args = ap.parse_args()
    peak = 6550.4
    try:
        import json
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    lines = ["# r02: the reference's published performance table re-measured on one B200", "",
             "Protocol of `/root/reference/test/test_performance.py:3-32` (32 MB buffer, out-of-place, 1 warm-up + 10 iterations,",
             "GFLOPS = 5e-9 * N log2 N * batch / t), shapes of `test_performance.py:40-44`; published pyfft numbers are the",
             "Tesla C2050 column of `doc/source/index.rst:352-374` (other hardware, for context only).  `GB/s` = passes x 2 x bytes / t;",
             "a 32 MB working set lives in the 126 MB L2, so the first block is not an HBM measurement -- the second one",
             "(2 GiB buffer) is.  `tools/published_table.py`.", ""]
    for mib, title in ((32, "32 MB buffer (the reference's protocol; L2-resident on B200)"), (2048, "2 GiB buffer (HBM-resident)")):
        lines += ["## " + title, "", "| shape | batch | sp ms | sp GFLOPS | sp GB/s | published pyfft sp (C2050) | dp ms | dp GFLOPS | dp GB/s | published pyfft dp |",
                  "|---|---|---|---|---|---|---|---|---|---|"]
        for shape in SHAPES:
            sp = measure(shape, np.complex64, mib)
            dp = measure(shape, np.complex128, mib)
            pub = PUBLISHED.get(shape, (None, None))
            label = "[" + ", ".join(str(v) for v in list(reversed(shape)) + [1] * (3 - len(shape))) + "]"
            lines.append("| %s | %d | %.4f | %.0f | %.0f%s | %s | %.4f | %.0f | %.0f%s | %s |" % (
                label, sp[0], sp[1], sp[2], sp[3], " (%.2f)" % (sp[3] / peak) if mib > 32 else "", pub[0] if pub[0] else "–",
                dp[1], dp[2], dp[3], " (%.2f)" % (dp[3] / peak) if mib > 32 else "", pub[1] if pub[1] else "–"))
            print(lines[-1], flush=True)
            torch.cuda.empty_cache()
        lines.append("")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
