"""GPU parity tests of the four-step (multi-pass) path: axes longer than one CTA can hold
(contiguous axis > 2^14 single / 2^13 double, strided axes > 2^11) are split into a transposing
pass with the inter-pass twiddle fused into its stores + the remaining axis (api.cu add_axis()).
This is the counterpart of the reference's global-kernel chains (pyfft/plan.py:141-143,
pyfft/kernel.py:259-283); the reference's own parity grid reaches 1D 2^20 (test/test_errors.py:125).
Checked against the float64 numpy.fft oracle within 1e-5*log2(N) / 1e-13*log2(N) relative L2 and the
reference's eps (1.1e-6 / 1e-11 on sum|a-b|/sum|a|), in-place == out-of-place bit for bit, round trips,
and against the restated reference algorithm."""
import ctypes

import numpy as np
import pytest

from oracle import numpy_oracle as no
from oracle import pyfft_restatement as pr
from test_parity_gpu import _gpu, _run

pytestmark = pytest.mark.gpu

# (shape, batch): long contiguous axis (2 and 3 passes), long strided axes, mixtures
CASES = [
    ((1 << 15,), 3), ((1 << 16,), 2), ((1 << 17,), 1), ((1 << 18,), 2), ((1 << 20,), 1), ((1 << 20,), 3),
    ((1 << 22,), 1), ((1 << 23,), 1), ((1 << 24,), 1),
    ((4096, 64), 2), ((8192, 16), 1), ((4096, 2), 3), ((64, 1 << 15), 2), ((4096, 1 << 15), 1),
    ((4096, 4, 8), 2), ((2, 8192, 16), 1), ((4096, 8, 4096), 1),
]
DTYPES = [np.complex64, np.float32, np.complex128, np.float64]


def _ids(v):
    return "x".join(map(str, v)) if isinstance(v, tuple) else str(v)


@pytest.mark.parametrize("shape,batch", CASES, ids=_ids)
@pytest.mark.parametrize("dtype", DTYPES, ids=["c64", "f32split", "c128", "f64split"])
def test_long_axes_vs_numpy(cuda_device, shape, batch, dtype):
    from pyfft_b200.cuda import Plan
    size = int(np.prod(shape))
    if size * batch > (1 << 24) and dtype is not np.complex64:
        pytest.skip("beyond the host oracle budget; the complex64 case covers this shape")
    data = no.make_input(shape, batch, dtype, seed=7 + size % 89)
    z = (data[0] + 1j * data[1]) if isinstance(data, tuple) else data
    plan = Plan(shape, dtype=dtype, normalize=True, wait_for_finish=True)
    assert any("fs=" in l for l in plan.passes), plan.passes
    tol, eps = no.tolerance(dtype, size), no.reference_epsilon(dtype)
    ref_fw = no.fft_oracle(z, shape, batch)
    fw_out = _run(plan, cuda_device, data, batch, False, inplace=False)
    fw_in = _run(plan, cuda_device, data, batch, False, inplace=True)
    assert np.array_equal(fw_out, fw_in), "in-place and out-of-place forward differ"
    assert no.rel_l2(fw_in, ref_fw) < tol
    assert no.pyfft_difference(ref_fw, fw_in, batch) < eps
    fw_data = (np.ascontiguousarray(fw_in.real).astype(dtype), np.ascontiguousarray(fw_in.imag).astype(dtype)) \
        if isinstance(data, tuple) else fw_in.astype(dtype)
    back_in = _run(plan, cuda_device, fw_data, batch, True, inplace=True)
    back_out = _run(plan, cuda_device, fw_data, batch, True, inplace=False)
    assert np.array_equal(back_in, back_out)
    assert no.rel_l2(back_in, z) < tol
    assert no.pyfft_difference(z, back_in, batch) < eps


@pytest.mark.parametrize("shape,batch", [((1 << 16,), 2), ((4096, 32), 1)], ids=_ids)
@pytest.mark.parametrize("dtype", [np.complex64, np.complex128], ids=["c64", "c128"])
@pytest.mark.parametrize("inverse,normalize,scale", [(False, True, 1.0), (True, True, 1.0), (True, False, 1.0),
                                                     (False, True, 3.0), (True, True, 3.0)])
def test_long_axes_vs_reference_restatement(cuda_device, shape, batch, dtype, inverse, normalize, scale):
    """Same inputs through the restated reference algorithm (its multi-pass global-kernel chain) and
    through the four-step CUDA path, all scaling modes (pyfft/kernel.py:23-37)."""
    from pyfft_b200.cuda import Plan
    size = int(np.prod(shape))
    data = no.make_input(shape, batch, dtype, seed=41)
    re, im = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
    pre, pim = pr.pyfft_execute(re, im, shape, batch, inverse, normalize, scale)
    plan = Plan(shape, dtype=dtype, normalize=normalize, scale=scale)
    got = _run(plan, cuda_device, data, batch, inverse, inplace=False)
    want64 = no.fft_oracle(re.astype(np.float64) + 1j * im.astype(np.float64), shape, batch, inverse, normalize, scale)
    tol = no.tolerance(dtype, size)
    assert no.rel_l2(got, pre + 1j * pim) < tol
    assert no.rel_l2(got, want64) < tol
    assert no.rel_l2(got, want64) < 2.0 * no.rel_l2(pre + 1j * pim, want64) + 1e-16


def test_known_answers_long(cuda_device):
    """ones(8192) round trip with fast_math on and off (test/test_functionality.py:102-115) at a length
    that needs the multi-pass path here too (2^16), plus delta -> ones."""
    import torch
    from pyfft_b200.cuda import Plan
    n = 1 << 16
    for fast_math in (True, False):
        plan = Plan(n, dtype=np.complex64, fast_math=fast_math)
        a = torch.ones(n, dtype=torch.complex64, device=cuda_device)
        plan.execute(a)
        out = a.cpu().numpy()
        assert abs(out[0] - n) < 1e-6 * n and np.abs(out[1:]).max() < 1e-6 * n
        plan.execute(a, inverse=True)
        assert np.abs(a.cpu().numpy() - 1).mean() < 1e-6
    d = torch.zeros(n, dtype=torch.complex128, device=cuda_device)
    d[1] = 1
    Plan(n, dtype=np.complex128).execute(d)
    want = np.exp(-2j * np.pi * np.arange(n) / n)
    assert np.abs(d.cpu().numpy() - want).max() < 1e-14


def test_workspace_contract(cuda_device):
    """C ABI: four-step plans report a data-sized workspace for in-place executes (none for an
    out-of-place execute whose long axis is the first pass) and refuse to run without it."""
    import torch
    from pyfft_b200 import _lib
    lib = _lib.load()
    n, batch = 1 << 16, 3

    def make(dims):
        h = ctypes.c_void_p()
        _lib.check(lib.b2fft_plan_create(ctypes.byref(h), 3, (ctypes.c_int64 * 3)(*dims), _lib.F32, _lib.INTERLEAVED,
                                         1, 1.0, 1, 0))
        return h

    h = make((n, 1, 1))
    need = ctypes.c_size_t(123)
    _lib.check(lib.b2fft_plan_workspace_bytes(h, batch, ctypes.byref(need)))
    assert need.value == n * batch * 8
    _lib.check(lib.b2fft_plan_workspace_bytes_ex(h, batch, 0, ctypes.byref(need)))
    assert need.value == 0
    a = torch.zeros(n * batch, dtype=torch.complex64, device=cuda_device)
    b = torch.zeros_like(a)
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.b2fft_execute(h, a.data_ptr(), None, b.data_ptr(), None, 0, batch, stream) == _lib.OK
    assert lib.b2fft_execute(h, a.data_ptr(), None, a.data_ptr(), None, 0, batch, stream) == _lib.E_INVALID
    assert "workspace" in _lib.last_error()
    ws = torch.zeros(n * batch, dtype=torch.complex64, device=cuda_device)
    _lib.check(lib.b2fft_plan_set_workspace(h, ws.data_ptr(), ws.numel() * 8))
    assert lib.b2fft_execute(h, a.data_ptr(), None, a.data_ptr(), None, 0, batch, stream) == _lib.OK
    torch.cuda.synchronize()
    lib.b2fft_plan_destroy(h)
    # long strided axis after an X pass: out-of-place needs the workspace as well
    h = make((64, 4096, 1))
    _lib.check(lib.b2fft_plan_workspace_bytes_ex(h, 2, 0, ctypes.byref(need)))
    assert need.value == 64 * 4096 * 2 * 8
    lib.b2fft_plan_destroy(h)
    # short axes: never
    h = make((4096, 1, 1))
    _lib.check(lib.b2fft_plan_workspace_bytes(h, 5, ctypes.byref(need)))
    assert need.value == 0
    lib.b2fft_plan_destroy(h)


def test_long_axis_unaligned_and_changing_batch(cuda_device):
    """8-byte (not 16-byte) aligned buffers take the non-TMA fallback kernels of the transposing pass;
    one plan serves changing batch sizes (workspace re-sized, pyfft/plan.py:184-192)."""
    import torch
    from pyfft_b200.cuda import Plan
    n = 1 << 20
    plan = Plan(n, dtype=np.complex64)
    for batch in (2, 1, 3):
        data = no.make_input((n,), batch, np.complex64, seed=batch)
        want = no.fft_oracle(data, (n,), batch)
        buf = torch.zeros(n * batch + 1, dtype=torch.complex64, device=cuda_device)
        view = buf[1:]
        view.copy_(_gpu(data.ravel(), cuda_device))
        plan.execute(view, batch=batch)
        assert no.rel_l2(view.cpu().numpy().reshape(data.shape), want) < no.tolerance(np.complex64, n)
        out = torch.zeros(n * batch + 1, dtype=torch.complex64, device=cuda_device)
        aligned = _gpu(data.ravel(), cuda_device)
        plan.execute(aligned, out[1:], batch=batch)
        assert no.rel_l2(out[1:].cpu().numpy().reshape(data.shape), want) < no.tolerance(np.complex64, n)
