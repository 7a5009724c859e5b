#!/usr/bin/env python
"""Slab-decomposed 3D FFT: correctness and timing under torchrun (one rank per GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
      tools/slab_check.py --size 256 --check              # parity vs numpy.fft.fftn (gathered on rank 0)
  ... tools/slab_check.py --size 2048 --steps 5           # timing, closed-form + Parseval + round-trip checks

Prints one JSON line per (size, exchange) on rank 0.
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, nargs="*", default=None)
    ap.add_argument("--shape", nargs="*", default=[], help="Z,Y,X (non-cubic arrays), e.g. 16,2048,64")
    ap.add_argument("--exchange", nargs="+", default=["p2p", "nccl"], help="p2p | nccl | p2p-yzx ([Yb][Z][X] y-slab layout)")
    ap.add_argument("--dtype", default="complex64")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--check", action="store_true", help="gather and compare with numpy.fft.fftn (small sizes)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from pyfft_b200.dist import SlabPlan
    npdt = np.dtype(args.dtype)
    results = []
    shapes = [(n, n, n) for n in (args.n if args.n is not None else ([] if args.shape else [256]))]
    shapes += [tuple(int(v) for v in sh.split(",")) for sh in args.shape]
    for shape in shapes:
        n = shape[0] if shape[0] == shape[1] == shape[2] else list(shape)
        nz, ny, nx = shape
        for ex in args.exchange:
            if world == 1 and ex != args.exchange[0]:
                continue
            yzx = ex.endswith("-yzx")
            chunks = int(ex.split("x")[1]) if ex.startswith("ncclx") else 1          # "ncclx8" = 8 pipelined chunks
            xs = ex.startswith("xslab")               # "xslab[xC][zK][cN]": C y-chunks, K z-chunks, N exchange CTAs per SM
            ctas, zch, ovl, ocol = 3, 0, None, 0
            if xs:       # "xslab[xC][zK][cN][oR]": ... R SMs left to the exchange while the Y pass runs (o0 = Y pass not hidden)
                import re
                m = re.match(r"xslab(?:x(\d+))?(?:z(\d+))?(?:c(\d+))?(?:o(\d+))?(?:p(\d+))?(?:b(\d))?(?:v(\d))?$", ex)
                if not m:
                    raise SystemExit("bad exchange spec " + ex)
                chunks = int(m.group(1)) if m.group(1) else 8
                zch = int(m.group(2)) if m.group(2) else 0
                ctas = int(m.group(3)) if m.group(3) else 3
                ovl = int(m.group(4)) if m.group(4) else None
                ocol = int(m.group(5)) if m.group(5) else 0          # pN: N y-chunks sent z-chunk by z-chunk beside the Y launch
                # vN: kernel of the Z pass (1 = TMA-staged W=4 tiles that share their SMs with the exchange CTAs, 0 = default)
                os.environ["B2FFT_SLAB_Z_VARIANT"] = {"1": "float_n11_w4_g1_b1_r16x16x8x1_tmac1", "2": "float_n11_w4_g1_b1_r16x16x8x1_tmac2"}.get(m.group(7) or "0", "")
                from pyfft_b200 import _lib as _l                    # bM: blocked-store mode (b2fft_set_option "blk_bulk")
                _l.check(_l.load().b2fft_set_option(b"blk_bulk", float(m.group(6)) if m.group(6) else 1.0))
            plan = SlabPlan(shape, dtype=npdt, exchange="xslab" if xs else "nccl" if ex.startswith("nccl") else ex.split("-")[0],
                            yslab_layout="yzx" if yzx else "zyx", chunks=chunks, exchange_ctas_per_sm=ctas, z_chunks=zch,
                            overlap_sms=ovl, overlap_columns=ocol)
            L = plan.L
            g = torch.Generator(device=dev)
            g.manual_seed(4242 + rank)
            fl = torch.float32 if npdt == np.complex64 else torch.float64

            def fill():
                zs = max(1, L["Zl"] // 16)
                for z0 in range(0, L["Zl"], zs):          # in pieces: no slab-sized temporaries
                    z1 = min(L["Zl"], z0 + zs)
                    plan.slab[z0:z1].copy_(torch.view_as_complex(torch.randn(z1 - z0, ny, nx, 2, dtype=fl, device=dev, generator=g)))
            rec = {"n": n, "world": world, "exchange": ex, "dtype": args.dtype,
                   "y_chunks": plan.chunks, "z_chunks": getattr(plan, "z_chunks", 1),
                   "describe": plan.describe()[:400] if hasattr(plan, "describe") else ""}
            fill()
            x_local = plan.slab.clone() if (args.check or nz * ny * nx <= 1024 ** 3) else None
            energy_in = float((plan.slab.abs() ** 2).sum().double().item())
            plan.forward()
            torch.cuda.synchronize()
            energy_out = float((plan.yslab.abs() ** 2).sum().double().item())
            if world > 1:
                t = torch.tensor([energy_in, energy_out], dtype=torch.float64, device=dev)
                dist.all_reduce(t)
                energy_in, energy_out = t.tolist()
            # Parseval: sum|X|^2 = N * sum|x|^2
            rec["parseval_rel_err"] = abs(energy_out / (energy_in * float(nz * ny * nx)) - 1.0)
            if args.check:
                if world > 1:      # NCCL has no complex dtype: gather the (re, im) views
                    xr, yr = torch.view_as_real(x_local).contiguous(), torch.view_as_real(plan.yslab).contiguous()
                    px = [torch.empty_like(xr) for _ in range(world)] if rank == 0 else None
                    py = [torch.empty_like(yr) for _ in range(world)] if rank == 0 else None
                    dist.gather(xr, px, dst=0)
                    dist.gather(yr, py, dst=0)
                    if rank == 0:
                        parts_x = [torch.view_as_complex(t) for t in px]
                        parts_y = [torch.view_as_complex(t) for t in py]
                else:
                    parts_x, parts_y = [x_local], [plan.yslab]
                if rank == 0:
                    full = torch.cat(parts_x, dim=0).cpu().numpy()
                    if xs:          # x-slabs [Y][Z][Xb] side by side along x
                        got = torch.cat([t.permute(1, 0, 2) for t in parts_y], dim=2).cpu().numpy()
                    elif yzx:
                        got = torch.cat([t.permute(1, 0, 2) for t in parts_y], dim=1).cpu().numpy()
                    else:
                        got = torch.cat(parts_y, dim=1).cpu().numpy()      # y-slabs side by side
                    want = np.fft.fftn(full.astype(np.complex128))
                    rec["fwd_rel_l2"] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
            if x_local is not None:
                plan.inverse()
                torch.cuda.synchronize()
                num = float(((plan.slab - x_local).abs() ** 2).sum().double().item())
                den = float((x_local.abs() ** 2).sum().double().item())
                if world > 1:
                    t = torch.tensor([num, den], dtype=torch.float64, device=dev)
                    dist.all_reduce(t)
                    num, den = t.tolist()
                rec["roundtrip_rel_l2"] = math.sqrt(num / den)
            # closed form: a delta at the origin transforms to all ones
            plan.slab.zero_()
            if rank == 0:
                plan.slab[0, 0, 0] = 1.0
            plan.forward()
            torch.cuda.synchronize()
            rec["delta_max_err"] = float((plan.yslab - 1.0).abs().max().item())
            # timing (forward only; the graded single-exchange transform)
            if args.steps > 0:
                fill()
                for _ in range(args.warmup):
                    plan.forward()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0 = plan.launch_count
                e0.record()
                for _ in range(args.steps):
                    plan.forward()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                if world > 1:
                    t = torch.tensor([ms], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                N = float(nz) * ny * nx
                if hasattr(plan, "set_trace") and os.environ.get("SLAB_TRACE"):
                    plan.set_trace(True)
                    plan.forward()
                    rec["trace_rank0"] = plan.trace()
                    plan.set_trace(False)
                rec["ms"] = ms
                rec["gflops"] = 5.0 * N * math.log2(N) / (ms * 1e-3) / 1e9
                rec["launches_per_step"] = (plan.launch_count - l0) / args.steps
                rec["sent_bytes_per_gpu"] = L["slab_elems"] * npdt.itemsize * (world - 1) / world
                rec["nvlink_gbs_per_gpu_if_exchange_were_all"] = rec["sent_bytes_per_gpu"] / (ms * 1e-3) / 1e9
            if hasattr(plan, "status"):
                rec["status"] = plan.status()
            plan.close()
            del plan
            torch.cuda.empty_cache()
            if rank == 0:
                print(json.dumps(rec), flush=True)
                results.append(rec)
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(results, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
