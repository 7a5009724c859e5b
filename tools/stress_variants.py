#!/usr/bin/env python
"""Runs kernel variants repeatedly on the same input and reports run-to-run differences
(race detector) plus the error against numpy.fft.  python tools/stress_variants.py --filter tma --reps 40"""
import argparse, ctypes, os, re, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filter", default="tma")
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--inplace", type=int, default=0)
    args = ap.parse_args()
    import torch
    from pyfft_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    buf = ctypes.create_string_buffer(256)
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(lib.b2fft_num_variants()):
        lib.b2fft_variant_info(i, buf, len(buf))
        f = buf.value.decode().split()
        name, prec, lg, W, G = f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4])
        if not re.search(args.filter, name) or W != 1:
            continue   # strided variants are stressed by tests/test_parity_gpu.py::test_persistent_tma_kernels_are_race_free
        n = 1 << lg
        rows = max(args.rows, 8 * G * 148)
        cdt = np.complex64 if prec == 0 else np.complex128
        rng = np.random.default_rng(1)
        x = (rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n))).astype(cdt)
        want = np.fft.fft(x.astype(np.complex128), axis=1)
        a = torch.from_numpy(x).to(dev)
        first = None
        nbad_runs = 0
        detail = ""
        for rep in range(args.reps):
            if args.inplace:
                b = a.clone()
                _lib.check(lib.b2fft_run_variant(i, b.data_ptr(), None, b.data_ptr(), None, 0, 0, rows, 1, 0, stream))
            else:
                b = torch.zeros_like(a)
                _lib.check(lib.b2fft_run_variant(i, a.data_ptr(), None, b.data_ptr(), None, 0, 0, rows, 1, 0, stream))
            out = b.cpu().numpy()
            if first is None:
                first = out
                err = np.linalg.norm(out - want) / np.linalg.norm(want)
            elif not np.array_equal(out, first):
                nbad_runs += 1
                if not detail:
                    badrows = np.where((out != first).any(axis=1))[0]
                    r = badrows[0]
                    # which of the two is wrong, and where do the INPUT-domain differences sit?
                    e_first = np.abs(first[r] - want[r]).max(); e_out = np.abs(out[r] - want[r]).max()
                    wrong = out[r] if e_out > e_first else first[r]
                    back = np.fft.ifft(wrong.astype(np.complex128))            # implied input of the wrong row
                    din = np.abs(back - x[r].astype(np.complex128))
                    idx = np.where(din > 1e-3 * np.abs(x[r]).max())[0]
                    detail = " rows=%s implied-input diffs at n=%s..%s (%d elems; head %s)" % (badrows[:6].tolist(), idx.min() if idx.size else -1, idx.max() if idx.size else -1, idx.size, idx[:12].tolist())
        print("%-46s rows=%d err=%.2e nondeterministic runs: %d/%d%s" % (name, rows, err, nbad_runs, args.reps - 1, detail), flush=True)

if __name__ == "__main__":
    main()
