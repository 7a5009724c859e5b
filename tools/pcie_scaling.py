#!/usr/bin/env python
"""Host <-> device copy bandwidth with 1..N ranks copying at the same time (torchrun, one rank per GPU).

Why: bench.py's end-to-end number (pinned host -> H2D -> FFT -> D2H) stopped scaling with the GPU count in round 1
(46 -> 8 GB/s per GPU each way from 1 to 8 GPUs).  This isolates the copies from the FFT: every rank moves the same
pinned buffers with plain cudaMemcpyAsync (torch copy_ non_blocking) -- H2D alone, D2H alone, both directions at
once -- and rank 0 prints per-rank and aggregate GB/s.  With --bind each rank first pins itself to the CPU cores
NVML reports as local to its GPU (so the pinned pages are first-touched on that NUMA node).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_scaling.py [--bind]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist


def gpu_cpu_affinity(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = (os.cpu_count() + 63) // 64
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [i * 64 + b for i, m in enumerate(masks) for b in range(64) if (m >> b) & 1]
        numa = None
        try:
            numa = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            pass
        return cpus, numa
    except Exception:
        return [], None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bind", action="store_true")
    ap.add_argument("--chunks", type=int, default=1, help="split every copy into this many cudaMemcpyAsync calls")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cpus, numa = gpu_cpu_affinity(local)
    if args.bind and cpus:
        try:
            os.sched_setaffinity(0, cpus)
        except OSError:
            pass
    n = args.mib << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(0)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    step = n // args.chunks

    def h2d(stream):
        with torch.cuda.stream(stream):
            for c in range(args.chunks):
                d_a[c * step:(c + 1) * step].copy_(h_in[c * step:(c + 1) * step], non_blocking=True)

    def d2h(stream):
        with torch.cuda.stream(stream):
            for c in range(args.chunks):
                h_out[c * step:(c + 1) * step].copy_(d_b[c * step:(c + 1) * step], non_blocking=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def timed(fn):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(args.reps):
            fn()
        torch.cuda.current_stream(dev).wait_stream(s1)
        torch.cuda.current_stream(dev).wait_stream(s2)
        e1.record()
        barrier()
        return n * args.reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    res = {"h2d": timed(lambda: h2d(s1)), "d2h": timed(lambda: d2h(s2)), "both_each_way": timed(lambda: (h2d(s1), d2h(s2)))}
    rec = {"rank": rank, "numa": numa, "cpus": "%d..%d (%d)" % (cpus[0], cpus[-1], len(cpus)) if cpus else None,
           **{k: round(v, 2) for k, v in res.items()}}
    allrec = [None] * world
    if world > 1:
        dist.all_gather_object(allrec, rec)
    else:
        allrec = [rec]
    if rank == 0:
        out = {"world": world, "bind": args.bind, "mib": args.mib, "chunks": args.chunks, "host_cpus": os.cpu_count(),
               "sum_h2d": round(sum(r["h2d"] for r in allrec), 1), "sum_d2h": round(sum(r["d2h"] for r in allrec), 1),
               "sum_both_each_way": round(sum(r["both_each_way"] for r in allrec), 1), "ranks": allrec}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
