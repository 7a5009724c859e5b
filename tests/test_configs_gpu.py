"""GPU parity tests at the shapes BASELINE.json names (cfg3, cfg4) and at cfg5's per-pass geometry.

cfg3  2D complex64 1024x1024 (strided Y pass of length 1024 at 8 KiB pitch), all four dtypes.
cfg4  3D complex128 256^3, fast_math off (true-division scaling), forward + inverse, normalize on/off,
      in place and out of place.
cfg5  3D complex64 2048^3 does not fit a host oracle (128 GiB in float64); its three passes are pinned
      one geometry at a time inside real plans: a 2048-long Y axis at 16 KiB pitch, a 2048-long Z axis
      at >= 256 KiB pitch (where the planner switches kernel variants, api.cu Registry::pick), and
      2048-long X rows, each against the float64 numpy.fft oracle.
Tolerances: relative L2 <= 1e-5*log2(N) single / 1e-13*log2(N) double (BASELINE.json north_star) and
the reference's own eps 1.1e-6 / 1e-11 on sum|a-b|/sum|a| (test/test_errors.py:20-23)."""
import numpy as np
import pytest

from oracle import numpy_oracle as no
from test_parity_gpu import _gpu, _run

pytestmark = pytest.mark.gpu


def _check_both_ways(plan, dev, shape, batch, dtype, data, normalize=True, scale=1.0):
    size = int(np.prod(shape))
    z = (data[0] + 1j * data[1]) if isinstance(data, tuple) else data
    tol, eps = no.tolerance(dtype, size), no.reference_epsilon(dtype)
    ref_fw = no.fft_oracle(z, shape, batch, False, normalize, scale)
    fw_out = _run(plan, dev, data, batch, False, inplace=False)
    fw_in = _run(plan, dev, data, batch, False, inplace=True)
    assert np.array_equal(fw_out, fw_in), "in-place and out-of-place forward differ"
    assert no.rel_l2(fw_in, ref_fw) < tol
    assert no.pyfft_difference(ref_fw, fw_in, batch) < eps
    fw_data = (np.ascontiguousarray(fw_in.real).astype(dtype), np.ascontiguousarray(fw_in.imag).astype(dtype)) \
        if isinstance(data, tuple) else fw_in.astype(dtype)
    ref_bw = no.fft_oracle(fw_in, shape, batch, True, normalize, scale)
    back_in = _run(plan, dev, fw_data, batch, True, inplace=True)
    back_out = _run(plan, dev, fw_data, batch, True, inplace=False)
    assert np.array_equal(back_in, back_out), "in-place and out-of-place inverse differ"
    assert no.rel_l2(back_in, ref_bw) < tol
    assert no.pyfft_difference(ref_bw, back_in, batch) < eps
    return fw_in, back_in


@pytest.mark.parametrize("dtype", [np.complex64, np.float32, np.complex128, np.float64],
                         ids=["c64", "f32split", "c128", "f64split"])
def test_cfg3_1024x1024_batch4(cuda_device, dtype):
    from pyfft_b200.cuda import Plan
    shape, batch = (1024, 1024), 4
    data = no.make_input(shape, batch, dtype, seed=1003)
    plan = Plan(shape, dtype=dtype, normalize=True)
    assert len(plan.passes) == 2, plan.passes
    _, back = _check_both_ways(plan, cuda_device, shape, batch, dtype, data)
    z = (data[0] + 1j * data[1]) if isinstance(data, tuple) else data
    assert no.rel_l2(back, z) < no.tolerance(dtype, 1 << 20)          # normalised round trip


@pytest.mark.parametrize("normalize", [True, False], ids=["norm", "nonorm"])
def test_cfg4_256cubed_c128_accurate_math(cuda_device, normalize):
    """BASELINE config 4: 3D complex128 256^3, fast_math off, forward + inverse, with and without normalisation."""
    from pyfft_b200.cuda import Plan
    shape = (256, 256, 256)
    data = no.make_input(shape, 1, np.complex128, seed=1004)
    plan = Plan(shape, dtype=np.complex128, normalize=normalize, fast_math=False)
    assert len(plan.passes) == 3, plan.passes
    _, back = _check_both_ways(plan, cuda_device, shape, 1, np.complex128, data, normalize=normalize)
    want = data if normalize else data * float(1 << 24)
    assert no.rel_l2(back, want) < no.tolerance(np.complex128, 1 << 24)


def test_cfg4_256cubed_split_f64(cuda_device):
    from pyfft_b200.cuda import Plan
    shape = (256, 256, 256)
    data = no.make_input(shape, 1, np.float64, seed=1005)
    plan = Plan(shape, dtype=np.float64, fast_math=False)
    _check_both_ways(plan, cuda_device, shape, 1, np.float64, data)


# numpy-order (z, y, x) shapes whose passes have cfg5's geometry
CFG5_LIKE = [
    ((8, 2048, 2048), "Y"),      # 2048-long Y axis at 16 KiB pitch (the TMA-staged strided kernel) + 2048-long rows
    ((2048, 2048, 8), "YZ"),     # 2048-long Y (64 B pitch) and Z (128 KiB pitch) axes over a narrow inner dimension
    ((2048, 64, 512), "Z"),      # 2048-long Z axis at 256 KiB pitch: the large-pitch variant choice of Registry::pick
]


@pytest.mark.parametrize("shape,axes", CFG5_LIKE, ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) else v)
def test_cfg5_pass_geometries(cuda_device, shape, axes):
    import torch
    from pyfft_b200.cuda import Plan
    data = no.make_input(shape, 1, np.complex64, seed=1005 + shape[1])
    plan = Plan(shape, dtype=np.complex64)
    passes = plan.passes
    assert len(passes) == 3 and all("fs=" not in p for p in passes), passes
    for ax in axes:
        assert any(p.startswith("axis=%s n=2048" % ax) for p in passes), passes
    size = int(np.prod(shape))
    tol, eps = no.tolerance(np.complex64, size), no.reference_epsilon(np.complex64)
    want = no.fft_oracle(data, shape, 1)
    a = _gpu(data, cuda_device)
    b = torch.empty_like(a)
    plan.execute(a, b)
    got = b.cpu().numpy()
    assert no.rel_l2(got, want) < tol
    assert no.pyfft_difference(want, got, 1) < eps
    plan.execute(a)                                       # in place == out of place, bit for bit
    assert torch.equal(a, b)
    plan.execute(a, inverse=True)
    assert no.rel_l2(a.cpu().numpy(), data) < tol
    del want, got


def test_cfg5_axis_by_axis_2048(cuda_device):
    """Each pass of the 2048^3 plan on its own (axis masks of b2fft_plan_create_ex), on a [64][2048][2048]
    block for X/Y and a [2048][16][2048] block for Z, against numpy.fft.fft along that axis."""
    import ctypes
    import torch
    from pyfft_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(77)

    def run(dims_zyx, mask, axis):
        zdim, ydim, xdim = dims_zyx
        x = (rng.standard_normal(dims_zyx, dtype=np.float32) + 1j * rng.standard_normal(dims_zyx, dtype=np.float32)).astype(np.complex64)
        h = ctypes.c_void_p()
        _lib.check(lib.b2fft_plan_create_ex(ctypes.byref(h), (ctypes.c_int64 * 3)(xdim, ydim, zdim), mask, _lib.F32,
                                            _lib.INTERLEAVED, 1, 1.0, 1, 0, 0.0, 0))
        a = _gpu(x, cuda_device)
        _lib.check(lib.b2fft_execute(h, a.data_ptr(), None, a.data_ptr(), None, 0, 1,
                                     torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        got = a.cpu().numpy()
        lib.b2fft_plan_destroy(h)
        # sample lines along `axis` instead of transforming 2 GiB in float64 on the host
        moved_in, moved_out = np.moveaxis(x, axis, -1), np.moveaxis(got, axis, -1)
        idx = rng.integers(0, moved_in.shape[0], 24), rng.integers(0, moved_in.shape[1], 24)
        lines_in = moved_in[idx[0], idx[1]].astype(np.complex128)
        lines_out = moved_out[idx[0], idx[1]]
        assert no.rel_l2(lines_out, np.fft.fft(lines_in, axis=-1)) < no.tolerance(np.complex64, 2048)

    run((32, 2048, 2048), _lib.AXIS_X, 2)
    run((32, 2048, 2048), _lib.AXIS_Y, 1)
    run((2048, 32, 2048), _lib.AXIS_Z, 0)     # 512 KiB pitch


def test_identity_plan_applies_scale(cuda_device):
    """A plan whose transformed axes all have length 1 is the identity times the scale factor (forward * scale,
    inverse / scale), like every other shape (pyfft/kernel.py:23-37)."""
    import torch
    from pyfft_b200.cuda import Plan
    a = torch.arange(6, dtype=torch.float32, device=cuda_device).to(torch.complex64) + 1j
    want = a.clone()
    plan = Plan(1, dtype=np.complex64, scale=2.0)
    b = torch.empty_like(a)
    plan.execute(a, b, batch=6)
    assert torch.equal(b, want * 2.0)
    plan.execute(b, inverse=True, batch=6)
    assert torch.equal(b, want)
    re, im = want.real.contiguous(), want.imag.contiguous()
    Plan(1, dtype=np.float32, scale=4.0).execute(re, im, batch=6)
    assert torch.equal(re, want.real * 4.0) and torch.equal(im, want.imag * 4.0)
