"""GPU tests of the multi-GPU paths.  With one visible GPU they exercise the single-rank code path
(destination-blocked Y-pass stores, z-slab -> y-slab relayout, inverse); with >= 2 GPUs they also
run tools/slab_check.py under torchrun on 2 ranks (P2P-fused and NCCL exchanges)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import numpy_oracle as no

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("layout", ["zyx", "yzx"])
@pytest.mark.parametrize("shape,dtype", [((32, 64, 128), np.complex64), ((64, 32, 16), np.complex128), ((256, 256, 256), np.complex64)])
def test_slab_plan_single_rank(cuda_device, shape, dtype, layout):
    import torch
    from pyfft_b200.dist import SlabPlan
    plan = SlabPlan(shape, dtype=dtype, exchange="p2p", yslab_layout=layout)
    x = no.make_input(shape, 1, dtype, seed=9)[0]
    plan.slab.copy_(torch.from_numpy(x).to(cuda_device))
    y = plan.forward()
    torch.cuda.synchronize()
    want = np.fft.fftn(x.astype(np.complex128))
    tol = no.tolerance(dtype, int(np.prod(shape)))
    got = y.cpu().numpy()
    if layout == "yzx":
        got = got.transpose(1, 0, 2)
    assert no.rel_l2(got, want) < tol
    back = plan.inverse()
    torch.cuda.synchronize()
    assert no.rel_l2(back.cpu().numpy(), x) < tol
    plan.close()


@pytest.mark.parametrize("chunks,z_chunks", [(1, 1), (4, 1), (4, 2), (8, 4)])
@pytest.mark.parametrize("shape,dtype", [((32, 64, 128), np.complex64), ((64, 32, 32), np.complex128), ((256, 256, 256), np.complex64),
                                         ((16, 8, 2048), np.complex64), ((4096, 8, 64), np.complex64)])
def test_xslab_plan_single_rank(cuda_device, shape, dtype, chunks, z_chunks):
    """The native slab plan (b2fft_slab_*, csrc/slab.cu) on one rank: Y pass per z-chunk, chunked X pass with
    destination-blocked stores through the two-level outer index, chunked Z pass on the [Y][Z][X] result (the last
    shape has a 4096-long Z axis: its sub-plan needs the plan-owned workspace); inverse with source-blocked loads."""
    import torch
    from pyfft_b200.dist import SlabPlan
    plan = SlabPlan(shape, dtype=dtype, exchange="xslab", chunks=chunks, z_chunks=z_chunks)
    assert plan.status() == 0 and "x-slab" in plan.describe()
    x = no.make_input(shape, 1, dtype, seed=19)[0]
    plan.slab.copy_(torch.from_numpy(x).to(cuda_device))
    y = plan.forward()
    torch.cuda.synchronize()
    want = np.fft.fftn(x.astype(np.complex128))
    tol = no.tolerance(dtype, int(np.prod(shape)))
    assert no.rel_l2(y.cpu().numpy().transpose(1, 0, 2), want) < tol
    back = plan.inverse()
    torch.cuda.synchronize()
    assert no.rel_l2(back.cpu().numpy(), x) < tol
    y2 = plan.forward()                      # a second round trip reuses the flag words / epochs
    torch.cuda.synchronize()
    assert no.rel_l2(y2.cpu().numpy().transpose(1, 0, 2), want) < tol
    assert plan.status() == 0 and plan.launch_count > 0
    plan.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("shape", [(64, 64, 2048), (16, 128, 4096), (128, 128, 128)])
def test_xslab_blocked_store_modes(cuda_device, shape, mode):
    """The three ways the exchange pass stores its destination-blocked lines (b2fft_set_option "blk_bulk": warp stores,
    bulk copies staged in the exchange buffer, bulk copies from their own staging buffer that drain while the CTA goes
    on) give the same x-slabs, bit for bit."""
    import torch
    from pyfft_b200 import _lib
    from pyfft_b200.dist import SlabPlan
    lib = _lib.load()
    x = no.make_input(shape, 1, np.complex64, seed=39)[0]
    outs = []
    try:
        for m in (1, mode):
            _lib.check(lib.b2fft_set_option(b"blk_bulk", float(m)))
            plan = SlabPlan(shape, dtype=np.complex64, exchange="xslab", chunks=4, z_chunks=2)
            plan.slab.copy_(torch.from_numpy(x).to(cuda_device))
            outs.append(plan.forward().clone())
            torch.cuda.synchronize()
            plan.close()
    finally:
        _lib.check(lib.b2fft_set_option(b"blk_bulk", 1.0))
    want = np.fft.fftn(x.astype(np.complex128))
    assert no.rel_l2(outs[1].cpu().numpy().transpose(1, 0, 2), want) < no.tolerance(np.complex64, int(np.prod(shape)))
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("chunks,columns", [(4, 0), (8, 0), (8, 5), (8, 8)])
def test_xslab_plan_hidden_y_pass_single_rank(cuda_device, chunks, columns):
    """Overlap mode of the native slab plan (b2fft_slab_plan_set_overlap): the Y pass of a 2048-long axis is ONE persistent
    launch (streamed fused two-step kernel) on all but 40 SMs that publishes a progress counter per z-chunk; the X
    passes of z-chunk k wait for counter k on another stream.  Same result as the event-ordered schedule, repeatedly
    (the counters are re-armed every call)."""
    import torch
    from pyfft_b200.dist import SlabPlan
    shape, dtype = (16, 2048, 64), np.complex64
    # chunks = 8: the launches that go out beside the Y pass cover 3 (or `columns`) of the 8 y-chunks -- a line count that is
    # not a power of two
    plan = SlabPlan(shape, dtype=dtype, exchange="xslab", chunks=chunks, z_chunks=4, overlap_sms=40, overlap_columns=columns)
    assert "hidden under the exchange" in plan.describe() and "_fused2p" in plan.describe()
    ref = SlabPlan(shape, dtype=dtype, exchange="xslab", chunks=chunks, z_chunks=4, overlap_sms=0)
    assert "hidden under the exchange" not in ref.describe()
    x = no.make_input(shape, 1, dtype, seed=29)[0]
    want = np.fft.fftn(x.astype(np.complex128))
    tol = no.tolerance(dtype, int(np.prod(shape)))
    for _ in range(3):
        plan.slab.copy_(torch.from_numpy(x).to(cuda_device))
        ref.slab.copy_(torch.from_numpy(x).to(cuda_device))
        y, yr = plan.forward(), ref.forward()
        torch.cuda.synchronize()
        assert no.rel_l2(y.cpu().numpy().transpose(1, 0, 2), want) < tol
        assert torch.equal(y, yr)                    # same kernels on the same data: bit-identical
        back = plan.inverse()
        torch.cuda.synchronize()
        assert no.rel_l2(back.cpu().numpy(), x) < tol
    assert plan.status() == 0
    plan.close()
    ref.close()


@pytest.mark.parametrize("shape,wave", [((64, 128, 32), (5, 17, 3)), ((16, 2048, 64), (15, 1029, 63)), ((8, 16, 2048), (1, 0, 2047))])
def test_xslab_plane_wave_closed_form(cuda_device, shape, wave):
    """Closed form used to verify transforms too large for a float64 oracle: a plane wave
    exp(2 pi i (kz z/Z + ky y/Y + kx x/X)) transforms to N at (kz, ky, kx) and zero elsewhere."""
    import torch
    from pyfft_b200.dist import SlabPlan
    Z, Y, X = shape
    kz, ky, kx = wave
    plan = SlabPlan(shape, dtype=np.complex64, exchange="xslab", chunks=4, z_chunks=2, normalize=False)
    z = torch.arange(Z, device=cuda_device, dtype=torch.float64).view(Z, 1, 1) * (kz / Z)
    y = torch.arange(Y, device=cuda_device, dtype=torch.float64).view(1, Y, 1) * (ky / Y)
    x = torch.arange(X, device=cuda_device, dtype=torch.float64).view(1, 1, X) * (kx / X)
    ph = 2.0 * np.pi * ((z + y + x) % 1.0)
    plan.slab.copy_(torch.complex(torch.cos(ph), torch.sin(ph)).to(torch.complex64))
    out = plan.forward()                                    # [Y][Z][X] on one rank
    torch.cuda.synchronize()
    n = float(Z * Y * X)
    want = torch.zeros_like(out)
    want[ky, kz, kx] = n
    err = float((out - want).abs().max().item()) / n
    assert err < 1e-5 * np.log2(n), err
    plan.close()


def test_batch_sharded_plan_single_rank(cuda_device):
    import torch
    from pyfft_b200.dist import BatchShardedPlan
    p = BatchShardedPlan((64, 64), dtype=np.complex64)
    data = no.make_input((64, 64), 5, np.complex64, seed=3)
    a = torch.from_numpy(data).to(cuda_device)
    p.execute(a, batch=5)
    assert no.rel_l2(a.cpu().numpy(), no.fft_oracle(data, (64, 64), 5)) < no.tolerance(np.complex64, 4096)


def test_slab_two_ranks_if_available(cuda_device):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "slab_check.py"), "--size", "64", "128",
           "--check", "--steps", "0", "--exchange", "p2p", "nccl", "p2p-yzx", "ncclx4", "xslab", "xslabx2", "xslabx4z2"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    recs = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(recs) == 14
    for r in recs:
        assert r["fwd_rel_l2"] < 1e-5 * 21 and r["roundtrip_rel_l2"] < 1e-5 * 21 and r["delta_max_err"] < 1e-4, r
    # the Y pass hidden under the exchange (progress counters instead of events), 2048-long Y axis, against the plain schedule
    cmd = cmd[:cmd.index("--size")] + ["--shape", "32,2048,64", "--check", "--steps", "0", "--exchange", "xslabx4z4o40", "xslabx4z4o0"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    recs = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(recs) == 2 and "hidden under the exchange" in recs[0]["describe"] and "hidden" not in recs[1]["describe"]
    for r in recs:
        assert r["fwd_rel_l2"] < 1e-5 * 22 and r["roundtrip_rel_l2"] < 1e-5 * 22 and r["delta_max_err"] < 1e-4 and r["status"] == 0, r
