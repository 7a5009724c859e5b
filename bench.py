#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched C2C FFT hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b2fft|reference] [--workload cfg2]

Metric (BASELINE.json): C2C FFT GFLOP/s = 5*N*log2(N)*batch / t, N = x*y*z (reference
test/test_performance.py:24), plus the achieved fraction of the HBM roofline.  A "step" is one
out-of-place forward execute of the whole batch (the reference times out-of-place executes,
test/test_performance.py:26-30).  Default workload = BASELINE.json configs[1]:
batched 1D complex64 N=4096 batch=65536, normalize=True.

One JSON line on stdout (rank 0).  Multi-GPU (torchrun, one rank per GPU): every rank transforms
its own full-size batch (weak scaling, no data-path collective), time = max over ranks.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "C2C FFT GFLOP/s (5N log2N)"

# name -> (numpy-order shape, batch, dtype, layout, passes P, description)
WORKLOADS = {
    "cfg1": ((1024,), 16, "complex64", 1, "1D complex64 N=1024 batch=16 interleaved"),
    "cfg2": ((4096,), 65536, "complex64", 1, "1D complex64 N=4096 batch=65536 interleaved normalize=True"),
    "cfg2s": ((4096,), 65536, "float32", 1, "1D split float32 re/im N=4096 batch=65536 normalize=True"),
    "cfg3": ((1024, 1024), 256, "complex64", 2, "2D complex64 1024x1024 batch=256"),
    "cfg4": ((256, 256, 256), 1, "complex128", 3, "3D complex128 256^3 fast_math off"),
    # one big transform: N=1 runs an ordinary in-place Plan, N>1 the x-slab decomposition (strong scaling)
    "cfg5": ((2048, 2048, 2048), 1, "complex64", 3, "3D complex64 2048^3, slab-decomposed over the ranks (x-slab exchange over NVLink)"),
    "cfg5s": ((512, 512, 512), 1, "complex64", 3, "3D complex64 512^3 slab-decomposed (small stand-in for cfg5)"),
}


CSIZE = {"complex64": 8, "float32": 8, "complex128": 16, "float64": 16}      # bytes per complex element


def flops(shape, batch):
    n = int(np.prod(shape))
    return 5.0 * n * math.log2(n) * batch


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(workload)
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock + throttle reasons while the timed region runs."""

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self._h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self._nv = pynvml
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            try:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None

    def run(self):
        if self._h is None:
            return
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_min_mhz": float(min(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "window": "all timed regions of this run (headline loop, e2e, per-config and slab records)"}


# ----------------------------------------------------------------------------------------- CPU arms
def host_threads():
    """All host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU
    arms are meant to use the whole host, so the thread count is passed to the C port explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_port_rate(shape, dtype, target_seconds=6.0, max_batch=65536):
    """Times oracle/pyfft_port.c (the restated reference algorithm, all host threads) on a bounded
    batch of the same transform; returns (GFLOP/s, threads, sample description, seconds)."""
    from oracle import numpy_oracle as no
    from oracle import pyfft_port as pp
    threads = host_threads()
    size = int(np.prod(shape))
    b = max(threads * 4, 64)
    b = min(b, max_batch)
    x = no.make_input(shape, b, dtype, seed=1)
    if isinstance(x, tuple):
        x = (x[0] + 1j * x[1]).astype(np.complex64 if x[0].dtype == np.float32 else np.complex128)
    pp.execute(x, shape, b, nthreads=threads)                                   # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    pp.execute(x, shape, b, nthreads=threads)
    dt = time.perf_counter() - t0
    b2 = int(min(max_batch, max(b, b * target_seconds / max(dt, 1e-6))))
    if b2 > b * 2:
        reps = b2 // b
        x = np.tile(x, (reps,) + (1,) * (x.ndim - 1))
        b = b * reps
        t0 = time.perf_counter()
        pp.execute(x, shape, b, nthreads=threads)
        dt = time.perf_counter() - t0
    rate = flops(shape, b) / dt / 1e9
    return rate, threads, "batch=%d of shape %s (%d elements), %.2f s wall on %d threads" % (
        b, "x".join(map(str, shape)), size * b, dt, threads), dt


def numpy_scipy_rates(shape, dtype, batch=256):
    """Informational: pocketfft via numpy (1 thread) and scipy (all workers) on a small sample."""
    from oracle import numpy_oracle as no
    out = {}
    x = no.make_input(shape, batch, dtype, seed=2)
    if isinstance(x, tuple):
        x = x[0] + 1j * x[1]
    axes = tuple(range(1, x.ndim))
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        np.fft.fftn(x, axes=axes)
        best = min(best, time.perf_counter() - t0)
    out["numpy_fft_1thread_gflops"] = round(flops(shape, batch) / best / 1e9, 2)
    try:
        import scipy.fft as sf
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            sf.fftn(x, axes=axes, workers=os.cpu_count())
            best = min(best, time.perf_counter() - t0)
        out["scipy_fft_allcores_gflops"] = round(flops(shape, batch) / best / 1e9, 2)
    except Exception:
        pass
    out["sample_batch"] = batch
    return out


def workload_config(name, per_gpu_batch):
    """The `config` object both arms print for a workload (same keys and values, so the driver can match them)."""
    shape, batch, dtype, passes, desc = WORKLOADS[name]
    return {"workload": name + ": " + desc, "shape": list(shape), "batch": batch, "per_gpu_batch": per_gpu_batch,
            "layout": "split" if dtype in ("float32", "float64") else "interleaved", "normalize": True, "out_of_place": True}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle/pyfft_port.c; the real pyfft needs
    Python 2 + Mako + PyCUDA/PyOpenCL and cannot run in this image), with all host threads, on the same workload and
    config as the GPU arm: every step is one out-of-place forward transform of the WHOLE batch.  Rank 0 only: a host
    CPU baseline does not multiply with --gpus, the other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape, batch, dtype, passes, desc = WORKLOADS[args.workload]
    from oracle import pyfft_port as pp
    from oracle import numpy_oracle as no
    threads = host_threads()
    cdt = np.complex64 if CSIZE[dtype] == 8 else np.complex128
    size = int(np.prod(shape))
    # full batch per step when the host can hold it (input + output); otherwise the largest batch that fits 1/4 of RAM
    try:
        avail = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 64 << 30
    b = int(max(1, min(batch, (avail // 4) // (2 * size * np.dtype(cdt).itemsize))))
    small = no.make_input(shape, min(b, 256), cdt, seed=1)
    x = np.tile(small, (-(-b // small.shape[0]),) + (1,) * (small.ndim - 1))[:b]
    out = np.empty_like(x)
    for _ in range(args.warmup):
        pp.execute(x, shape, b, out=out, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pp.execute(x, shape, b, out=out, nthreads=threads)
    el = time.perf_counter() - t0
    rate = flops(shape, b) * args.steps / el / 1e9
    sample = "each step = batch %d of %d of shape %s, pyfft's algorithm restated in C (oracle/pyfft_port.c), OpenMP over lines, %d threads" % (
        b, batch, "x".join(map(str, shape)), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(rate, 3), "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(el / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if cdt == np.complex64 else "f64",
        "data": "synthetic", "config": workload_config(args.workload, batch),
        "cpu_baseline": {"value": round(rate, 3), "unit": "GFLOP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(rate, 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "host-CPU arm: one process on rank 0 whatever --gpus is; its value does not scale with the GPU count",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- per-pass / per-config records
def _alloc_random(n_el, dtype, dev, gen):
    """Device-resident standard-normal input: one complex tensor, or a (re, im) pair for the split layouts."""
    import torch
    tdt = {"complex64": torch.complex64, "complex128": torch.complex128, "float32": torch.float32, "float64": torch.float64}[dtype]
    if dtype in ("float32", "float64"):
        return (torch.randn(n_el, dtype=tdt, device=dev, generator=gen), torch.randn(n_el, dtype=tdt, device=dev, generator=gen))
    out = torch.empty(n_el, dtype=tdt, device=dev)
    outr = torch.view_as_real(out)
    step = max(1, n_el // 16)
    for i in range(0, n_el, step):                      # in pieces: no second full-size temporary
        outr[i:i + step].normal_(generator=gen)
    return out


def time_passes(shape, dtype, batch, dev, steps=5, warmup=3, data=None):
    """Every pass of the plan for `shape` timed on its own: one single-axis plan per axis (the axis mask of
    b2fft_plan_create_ex), in place on device-resident data, CUDA events on the launching stream.  Axes that need
    several launches (four-step split) report the launch count.  Returns a list of
    {axis, variant, launches, ms, gbs, frac} with frac = algorithmic bytes / time / measured HBM peak."""
    import ctypes
    import torch
    from pyfft_b200 import _lib
    lib = _lib.load()
    peak, _ = measured_peaks()
    dims = [1, 1, 1]
    for i, n in enumerate(reversed(shape)):
        dims[i] = int(n)
    split = dtype in ("float32", "float64")
    prec = _lib.F32 if CSIZE[dtype] == 8 else _lib.F64
    n_el = int(np.prod(shape)) * batch
    if data is None:
        g = torch.Generator(device=dev)
        g.manual_seed(99)
        data = _alloc_random(n_el, dtype, dev, g)
    ptrs = (data[0].data_ptr(), data[1].data_ptr()) if split else (data.data_ptr(), None)
    stream = torch.cuda.current_stream(dev).cuda_stream
    out = []
    for a, name in enumerate("XYZ"):
        if dims[a] <= 1:
            continue
        h = ctypes.c_void_p()
        _lib.check(lib.b2fft_plan_create_ex(ctypes.byref(h), (ctypes.c_int64 * 3)(*dims), 1 << a, prec,
                                            _lib.SPLIT if split else _lib.INTERLEAVED, 1, 1.0, 1, dev.index or 0, 0.0, 0))
        buf = ctypes.create_string_buffer(4096)
        lib.b2fft_plan_describe(h, buf, len(buf))
        lines = [l for l in buf.value.decode().splitlines() if l]
        need = ctypes.c_size_t(0)
        lib.b2fft_plan_workspace_bytes(h, batch, ctypes.byref(need))
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev) if need.value else None
        if ws is not None:
            lib.b2fft_plan_set_workspace(h, ws.data_ptr(), need.value)

        def run():
            _lib.check(lib.b2fft_execute(h, ptrs[0], ptrs[1], ptrs[0], ptrs[1], 0, batch, stream))
        for _ in range(warmup):
            run()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        alg = 2.0 * CSIZE[dtype] * n_el                                  # one read + one write of every element
        out.append({"axis": name, "variant": lines[-1].split("variant=")[1].split()[0] if lines else None,
                    "launches": len(lines), "ms": round(ms, 4), "gbs": round(alg / ms / 1e6, 1),
                    "frac": round(alg / ms / 1e6 / peak, 4)})
        lib.b2fft_plan_destroy(h)
        del ws
        for t in (data if split else (data,)):                           # undo the sqrt(N) growth of every unnormalised repeat
            t.mul_(float(dims[a]) ** (-(warmup + steps) / 2.0))
    return out


def config_record(name, dev, steps=10, warmup=3):
    """One BASELINE config measured the same way as the headline: out-of-place forward executes of the whole batch,
    device-resident, plus its per-pass table.  `frac` is the slowest pass's fraction (the dominant kernel)."""
    import torch
    from pyfft_b200.cuda import Plan
    shape, batch, dtype, passes, desc = WORKLOADS[name]
    split = dtype in ("float32", "float64")
    n_el = int(np.prod(shape)) * batch
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + sorted(WORKLOADS).index(name))
    a = _alloc_random(n_el, dtype, dev, g)
    b = (torch.empty_like(a[0]), torch.empty_like(a[1])) if split else torch.empty_like(a)
    plan = Plan(shape, dtype=np.dtype(dtype), normalize=True, fast_math=(name != "cfg4"),
                stream=torch.cuda.current_stream(dev), wait_for_finish=False)

    def step():
        if split:
            plan.execute(a[0], a[1], b[0], b[1], batch=batch)
        else:
            plan.execute(a, b, batch=batch)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = plan.launch_count
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    launches = (plan.launch_count - l0) // steps
    del b
    per_pass = time_passes(shape, dtype, batch, dev, steps=max(3, steps // 2), data=a)
    peak, _ = measured_peaks()
    alg = passes * 2.0 * CSIZE[dtype] * n_el
    rec = {"workload": desc, "ms": round(ms, 5), "gflops": round(flops(shape, batch) / ms / 1e6, 1), "launches_per_step": launches,
           "algorithmic_passes": passes, "hbm_frac_whole_step": round(alg / ms / 1e6 / peak, 4),
           "frac": min(q["frac"] for q in per_pass) if per_pass else None, "per_pass": per_pass}
    if 2 * CSIZE[dtype] * n_el < 2 * 126e6:
        rec["note"] = "working set fits the 126 MB L2: launch-latency-bound, not graded against the HBM roofline"
    return rec


def slab_record(args, world, rank, dev):
    """north_star's only collective path, at every N: ONE 2048^3 complex64 forward transform -- N=1: Plan.execute in
    place, N>1: SlabPlan (z-slabs in, x-slab exchange fused into the X pass over NVLink, x-slabs out) -- plus a 512^3
    run of the same code path gathered on rank 0 and compared with numpy.fft.fftn, and the single-GPU time measured in
    the same process so the parallel efficiency t1 / (N * tN) is self-contained."""
    import torch
    import torch.distributed as dist
    from pyfft_b200.cuda import Plan
    from pyfft_b200.dist import SlabPlan
    peak, _ = measured_peaks()
    rec = {"workload": WORKLOADS["cfg5"][4], "n_gpus": world}

    def fill(plan, n, gen):
        zl = plan.L["Zl"]
        zs = max(1, zl // 16)
        for z0 in range(0, zl, zs):
            z1 = min(zl, z0 + zs)
            plan.slab[z0:z1].copy_(torch.view_as_complex(torch.randn(z1 - z0, n, n, 2, device=dev, generator=gen)))

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- parity: 512^3 through the same code path, gathered on rank 0, against numpy.fft.fftn (float64)
    n = args.slab_parity_size
    gen = torch.Generator(device=dev)
    gen.manual_seed(5000 + rank)
    try:
        if world == 1:
            x = torch.view_as_complex(torch.randn(n, n, n, 2, device=dev, generator=gen))
            y = x.clone()
            plan = Plan((n, n, n), dtype=np.complex64, stream=torch.cuda.current_stream(dev), wait_for_finish=False)
            plan.execute(y)
            torch.cuda.synchronize(dev)
            full, got = x.cpu().numpy(), y.cpu().numpy()
            plan.execute(y, inverse=True)
            torch.cuda.synchronize(dev)
            num, den = float(((y - x).abs() ** 2).sum().double().item()), float((x.abs() ** 2).sum().double().item())
            del x, y, plan
        else:
            plan = SlabPlan((n, n, n), dtype=np.complex64, exchange="xslab", chunks=args.slab_chunks)
            fill(plan, n, gen)
            x_local = plan.slab.clone()
            plan.forward()
            torch.cuda.synchronize(dev)
            xr, yr = torch.view_as_real(x_local).contiguous(), torch.view_as_real(plan.xslab).contiguous()
            px = [torch.empty_like(xr) for _ in range(world)] if rank == 0 else None
            py = [torch.empty_like(yr) for _ in range(world)] if rank == 0 else None
            dist.gather(xr, px, dst=0)
            dist.gather(yr, py, dst=0)
            if rank == 0:
                full = torch.cat([torch.view_as_complex(t) for t in px], dim=0).cpu().numpy()
                got = torch.cat([torch.view_as_complex(t).permute(1, 0, 2) for t in py], dim=2).cpu().numpy()   # [Y][Z][Xb] side by side
            del px, py
            plan.inverse()
            torch.cuda.synchronize(dev)
            t = torch.tensor([float(((plan.slab - x_local).abs() ** 2).sum().double().item()),
                              float((x_local.abs() ** 2).sum().double().item())], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            num, den = t.tolist()
            plan.close()
            del plan, x_local
        parity = {"size": "%d^3" % n, "roundtrip_rel_l2": math.sqrt(num / den), "tolerance": 1e-5 * math.log2(float(n) ** 3),
                  "oracle": "numpy.fft.fftn in float64 on the gathered input (rank 0)"}
        if rank == 0:
            try:
                import scipy.fft as sf
                want = sf.fftn(full.astype(np.complex128), workers=host_threads())
            except Exception:
                want = np.fft.fftn(full.astype(np.complex128))
            parity["fwd_rel_l2"] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
            del want, full, got
        rec["parity"] = parity
    except Exception as exc:
        rec["parity"] = {"error": str(exc)[:300]}
    torch.cuda.empty_cache()

    # ---- parity of the pipeline the 2048^3 timing below runs from 4 ranks on (Y pass as ONE persistent launch with progress
    #      counters, hidden under the exchange): the 512^3 case above cannot take it (its Y axis is 512 long), so a
    #      [8G][2048][16G] array goes through the same default SlabPlan and is gathered on rank 0 against numpy.fft.fftn
    if world > 1:
        try:
            shp = (8 * world, 2048, 16 * world)
            plan = SlabPlan(shp, dtype=np.complex64, exchange="xslab", chunks=args.slab_chunks)
            hidden = "hidden under the exchange" in plan.describe()
            gen.manual_seed(5500 + rank)
            plan.slab.copy_(torch.view_as_complex(torch.randn(plan.L["Zl"], shp[1], shp[2], 2, device=dev, generator=gen)))
            x_local = plan.slab.clone()
            plan.forward()
            torch.cuda.synchronize(dev)
            xr, yr = torch.view_as_real(x_local).contiguous(), torch.view_as_real(plan.xslab).contiguous()
            px = [torch.empty_like(xr) for _ in range(world)] if rank == 0 else None
            py = [torch.empty_like(yr) for _ in range(world)] if rank == 0 else None
            dist.gather(xr, px, dst=0)
            dist.gather(yr, py, dst=0)
            plan.inverse()
            torch.cuda.synchronize(dev)
            t = torch.tensor([float(((plan.slab - x_local).abs() ** 2).sum().double().item()),
                              float((x_local.abs() ** 2).sum().double().item()), float(plan.status())], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            num, den, bad = t.tolist()
            ph = {"shape": list(shp), "hidden_y_pass": bool(hidden), "roundtrip_rel_l2": math.sqrt(num / den), "flag_timeouts": int(bad),
                  "tolerance": 1e-5 * math.log2(float(shp[0]) * shp[1] * shp[2])}
            if rank == 0:
                full = torch.cat([torch.view_as_complex(q) for q in px], dim=0).cpu().numpy()
                got = torch.cat([torch.view_as_complex(q).permute(1, 0, 2) for q in py], dim=2).cpu().numpy()
                want = np.fft.fftn(full.astype(np.complex128))
                ph["fwd_rel_l2"] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
                del full, got, want
            rec["parity_hidden_y"] = ph
            plan.close()
            del plan, x_local, px, py
        except Exception as exc:
            rec["parity_hidden_y"] = {"error": str(exc)[:300]}
        torch.cuda.empty_cache()

    # ---- timing at 2048^3
    n = args.slab_size
    size = float(n) ** 3
    steps = args.slab_steps
    try:
        gen.manual_seed(6000 + rank)
        if world == 1:
            a = _alloc_random(n ** 3, "complex64", dev, gen)
            plan = Plan((n, n, n), dtype=np.complex64, stream=torch.cuda.current_stream(dev), wait_for_finish=False)
            step = lambda: plan.execute(a)
            rec["plan"] = plan.passes
        else:
            plan = SlabPlan((n, n, n), dtype=np.complex64, exchange="xslab", chunks=args.slab_chunks)
            fill(plan, n, gen)
            step = plan.forward
            rec["plan"] = plan.describe() if hasattr(plan, "describe") else "x-slab"
        for _ in range(3):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = plan.launch_count
        barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        ms = allmax(e0.elapsed_time(e1) / steps)
        rec.update({"ms": round(ms, 3), "steps": steps, "gflops": round(5.0 * size * math.log2(size) / ms / 1e6, 1),
                    "gpu_launches_per_step": (plan.launch_count - l0) // steps})
        per_gpu_alg = 3 * 2.0 * 8 * size / world                          # P = 3 passes over this rank's share
        rec["hbm_frac"] = round(per_gpu_alg / ms / 1e6 / peak, 4)
        if world > 1:
            sent = 8.0 * size / world * (world - 1) / world
            rec["nvlink"] = {"sent_bytes_per_gpu": int(sent), "achieved_gbs": round(sent / ms / 1e6, 1),
                             "frac_of_770": round(sent / ms / 1e6 / 770.0, 4),
                             "note": "bytes each GPU sends over NVLink / WHOLE step time (local passes included)"}
            plan.close()
        else:
            rec["per_pass"] = time_passes((n, n, n), "complex64", 1, dev, steps=3, warmup=1, data=a)
            rec["frac"] = min(q["frac"] for q in rec["per_pass"])
            del a
        del plan
        torch.cuda.empty_cache()
        # ---- the single-GPU time beside it (rank 0 runs the ordinary in-place Plan on a full 2048^3 array)
        if world > 1:
            t1 = 0.0
            if rank == 0:
                a = _alloc_random(n ** 3, "complex64", dev, gen)
                p1 = Plan((n, n, n), dtype=np.complex64, stream=torch.cuda.current_stream(dev), wait_for_finish=False)
                for _ in range(2):
                    p1.execute(a)
                torch.cuda.synchronize(dev)
                e0.record()
                for _ in range(steps):
                    p1.execute(a)
                e1.record()
                torch.cuda.synchronize(dev)
                t1 = e0.elapsed_time(e1) / steps
                del a, p1
                torch.cuda.empty_cache()
            t1 = allmax(t1)
            rec["t1_ms"] = round(t1, 3)
            rec["parallel_efficiency"] = round(t1 / (world * ms), 4)
        else:
            rec["t1_ms"] = rec["ms"]
            rec["parallel_efficiency"] = 1.0
    except Exception as exc:
        rec["error"] = str(exc)[:300]
    return rec

# ----------------------------------------------------------------------------------------- GPU arm
def run_slab(args):
    """cfg5: ONE 3D transform.  N=1: Plan.execute in place; N>1: pyfft_b200.dist.SlabPlan (x-slab exchange).
    Strong scaling: value = 5 N log2 N / t with t the max over ranks."""
    import torch
    import torch.distributed as dist
    from pyfft_b200.cuda import Plan
    from pyfft_b200.dist import SlabPlan
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    shape, _, dtype, passes, desc = WORKLOADS[args.workload]
    n = shape[0]
    size = float(np.prod(shape))
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    launches = [0]
    if world == 1:
        a = torch.empty(shape, dtype=torch.complex64, device=dev)
        ar = torch.view_as_real(a)
        for z in range(0, n, 64):
            ar[z:z + 64].normal_(generator=g)
        plan = Plan(shape, dtype=np.complex64, stream=torch.cuda.current_stream(dev), wait_for_finish=False)
        desc_plan = plan.passes

        def step():
            plan.execute(a)
        count = lambda: plan.launch_count
    else:
        plan = SlabPlan(shape, dtype=np.complex64, exchange="xslab", chunks=8, exchange_ctas_per_sm=3)
        zs = max(1, plan.L["Zl"] // 16)
        for z0 in range(0, plan.L["Zl"], zs):
            z1 = min(plan.L["Zl"], z0 + zs)
            plan.slab[z0:z1].copy_(torch.view_as_complex(torch.randn(z1 - z0, n, n, 2, device=dev, generator=g)))
        desc_plan = ["x-slab: Y pass, 8 x (X pass with NVLink-blocked stores | barrier | Z pass)"]

        def step():
            plan.forward()
        count = lambda: plan.launch_count

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms /= args.steps
    value = 5.0 * size * math.log2(size) / (ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    per_gpu_bytes = 2.0 * 8 * size / world                 # one pass over this rank's share: read + write
    achieved = passes * per_gpu_bytes / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(per_gpu_bytes), "launches_per_step": passes,
                "note": "whole step (3 passes over this rank's share) / step time; at N>1 the step is NVLink-bound"}
    if world > 1:
        sent = 8.0 * size / world * (world - 1) / world
        roofline["nvlink"] = {"sent_bytes_per_gpu": int(sent), "achieved_gbs_if_step_were_all_exchange": round(sent / (ms * 1e-3) / 1e9, 1),
                              "peak_gbs": 770.0, "peak_source": "B200_PROFILING.md measured peer copy per direction"}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": round(value, 2), "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc, "in_place": True, "l2": "data exceeds the 126 MB L2; no flush needed",
                       "plan": desc_plan, "parallelism": "z-slabs over %d GPU(s), output x-slab distributed" % world if world > 1 else "single GPU"},
            "roofline": roofline, "cpu_baseline": None, "e2e": None, "gpu_launches": int(count() - l0), "clocks": clocks}), flush=True)
    if world > 1:
        plan.close()
        dist.destroy_process_group()


def run_b2fft(args):
    import torch
    import torch.distributed as dist
    from pyfft_b200.cuda import Plan
    from pyfft_b200.host import HostPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    shape, batch, dtype, passes, desc = WORKLOADS[args.workload]
    npdt = np.dtype(dtype)
    split = npdt.kind == "f"
    size = int(np.prod(shape))
    tdt = {"complex64": torch.complex64, "complex128": torch.complex128, "float32": torch.float32,
           "float64": torch.float64}[dtype]
    csize = {"complex64": 8, "float32": 8, "complex128": 16, "float64": 16}[dtype]   # bytes per complex element
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)

    def randn(n):
        if split:
            return torch.randn(n, dtype=tdt, device=dev, generator=g)
        fl = torch.float32 if tdt == torch.complex64 else torch.float64
        return torch.view_as_complex(torch.randn(n, 2, dtype=fl, device=dev, generator=g))

    n_el = size * batch
    fast_math = args.workload != "cfg4"
    plan = Plan(shape, dtype=npdt, normalize=True, fast_math=fast_math, stream=torch.cuda.current_stream(dev),
                wait_for_finish=False)
    if split:
        a_re, a_im, b_re, b_im = randn(n_el), randn(n_el), torch.empty(n_el, dtype=tdt, device=dev), torch.empty(
            n_el, dtype=tdt, device=dev)

        def step():
            plan.execute(a_re, a_im, b_re, b_im, batch=batch)
    else:
        a, b = randn(n_el), torch.empty(n_el, dtype=tdt, device=dev)

        def step():
            plan.execute(a, b, batch=batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = plan.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = plan.launch_count - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = flops(shape, batch) * world / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: algorithmic bytes per launch / average launch time
    peak, peak_src = measured_peaks()
    compulsory = 2.0 * csize * n_el                         # read every input once + write every output once
    alg_bytes_per_launch = compulsory                        # every pass is one full read + write of the data
    # the dominant kernel = the slowest pass.  One-pass plans: the step IS the kernel (events around the timed loop);
    # multi-pass plans: every pass is timed on its own (single-axis plans on the same data) and the slowest one is reported
    per_pass = None
    if passes > 1 and not split:
        per_pass = time_passes(shape, dtype, batch, dev, steps=max(3, min(args.steps, 10)), data=a)
    if per_pass:
        worst = min(per_pass, key=lambda q: q["frac"])
        per_launch_ms, dominant = worst["ms"], "%s pass (%s)" % (worst["axis"], worst["variant"])
    else:
        per_launch_ms, dominant = ms_per_step / passes, "the single pass" if passes == 1 else "average over the passes"
    achieved = alg_bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic(args.workload),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes_per_launch),
                "launches_per_step": passes, "dominant_kernel": dominant, "per_pass": per_pass,
                "whole_step_frac": round(passes * alg_bytes_per_launch / (ms_per_step * 1e-3) / 1e9 / peak, 4),
                "compulsory_fraction": round(compulsory / (ms_per_step * 1e-3) / 1e9 / peak, 4)}

    # ---- end to end: pinned host buffers, H2D + execute + D2H inside the timed region
    e2e = None
    if not split and not args.no_e2e:
        try:
            e2e_batch = batch
            pipe = HostPipeline(shape, dtype=npdt, batch=e2e_batch, chunks=args.e2e_chunks, slots=3, device=local_rank,
                                normalize=True, fast_math=fast_math)
            h_in = torch.empty(n_el, dtype=tdt).pin_memory()
            h_out = torch.empty(n_el, dtype=tdt).pin_memory()
            h_in.copy_(a)
            for _ in range(2):
                pipe.run(h_in, h_out)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps_e = max(3, min(args.steps, args.e2e_steps))
            l0 = pipe.launch_count
            barrier()
            e0.record()
            for _ in range(steps_e):
                pipe.run(h_in, h_out)
            e1.record()
            barrier()
            e_ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e_ms = float(t.item())
            e_val = flops(shape, e2e_batch) * world / (e_ms / steps_e * 1e-3) / 1e9
            e2e = {"value": round(e_val, 2), "unit": "GFLOP/s", "h2d_bytes_per_step": int(pipe.h2d_bytes),
                   "d2h_bytes_per_step": int(pipe.d2h_bytes), "ms_per_step": round(e_ms / steps_e, 3),
                   "steps": steps_e, "chunks": pipe.chunks, "gpu_launches": pipe.launch_count - l0,
                   "pcie_gbs_each_way": round(pipe.h2d_bytes / (e_ms / steps_e * 1e-3) / 1e9, 2)}
            # the result read back equals the device-resident result (same kernels, same data)
            if rank == 0:
                chk = torch.empty(size * 4, dtype=tdt, device=dev)
                Plan(shape, dtype=npdt, normalize=True, fast_math=fast_math).execute(a[:size * 4], chk, batch=4)
                if not torch.equal(chk.cpu(), h_out[:size * 4]):
                    e2e["warning"] = "host result differs from device result"
            del h_in, h_out, pipe
        except Exception as exc:   # keep the device-resident number even if pinned allocation fails
            e2e = {"value": None, "unit": "GFLOP/s", "error": str(exc)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cdt = np.complex64 if csize == 8 else np.complex128
            rate, threads, sample, _ = cpu_port_rate(shape, cdt, target_seconds=args.cpu_seconds)
            cpu_baseline = {"value": round(rate, 3), "unit": "GFLOP/s", "cores": threads, "kind": "port",
                            "sample": sample}
            cpu_baseline.update(numpy_scipy_rates(shape, cdt, batch=max(1, min(256, (1 << 22) // size))))
            cpu_baseline["host_cores"] = os.cpu_count()
        except Exception as exc:
            cpu_baseline = {"value": None, "error": str(exc)[:200]}

    # ---- every other BASELINE config, measured in this same run (rank 0's GPU; the other ranks wait)
    configs = None
    plan_desc = plan.passes
    if args.workload == "cfg2" and not args.no_configs:
        del plan
        if split:
            del a_re, a_im, b_re, b_im
        else:
            del a, b
        torch.cuda.empty_cache()
        if rank == 0:
            configs = {}
            for name in ("cfg1", "cfg2s", "cfg3", "cfg4"):
                try:
                    configs[name] = config_record(name, dev)
                except Exception as exc:
                    configs[name] = {"error": str(exc)[:200]}
                torch.cuda.empty_cache()
        barrier()
    # ---- the slab-decomposed 2048^3 transform (the path with a collective), at every N
    slab = None
    if args.workload == "cfg2" and not args.no_slab:
        slab = slab_record(args, world, rank, dev)
    clocks = sampler.stop()

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if csize == 8 else "f64", "data": "synthetic",
            "config": workload_config(args.workload, batch),
            "run": {"l2": "inputs+outputs (%.0f MiB) exceed the 126 MB L2; no flush needed" % (2 * csize * n_el / 2 ** 20)
                    if 2 * csize * n_el > 2 * 126e6 else "working set fits L2 (launch-bound config)",
                    "plan": plan_desc, "parallelism": "batch sharded over %d GPU(s), no collective" % world},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "configs": configs, "slab": slab,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b2fft", choices=["b2fft", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1/cfg2s/cfg3/cfg4 sub-records")
    ap.add_argument("--no-slab", action="store_true", help="skip the 2048^3 slab record")
    ap.add_argument("--slab-size", type=int, default=2048)
    ap.add_argument("--slab-parity-size", type=int, default=512)
    ap.add_argument("--slab-steps", type=int, default=5)
    ap.add_argument("--slab-chunks", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunks", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("cfg5", "cfg5s"):
        run_slab(args)
    else:
        run_b2fft(args)


if __name__ == "__main__":
    main()
