"""GPU parity tests: the CUDA path (through Plan -> ctypes -> C ABI) against
  * the float64 numpy.fft oracle, within the north_star tolerance 1e-5*log2(N) (single) /
    1e-13*log2(N) (double) relative L2, and the reference's own eps (1.1e-6 / 1e-11 on
    sum|a-b|/sum|a|, test/test_errors.py:20-23),
  * the numpy restatement of the reference's algorithm (oracle/pyfft_restatement.py),
  * the committed golden fixtures (tests/golden/*.npz),
on the reference's shape/batch/dtype grid (test/test_errors.py:122-140), in-place and
out-of-place, forward and inverse."""
import glob
import os

import numpy as np
import pytest

from oracle import numpy_oracle as no
from oracle import pyfft_restatement as pr

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _gpu(arr, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).to(dev)


def _run(plan, dev, data, batch, inverse, inplace):
    """Returns the complex result as a numpy array; data is complex (interleaved) or (re, im)."""
    import torch
    if isinstance(data, tuple):
        re, im = _gpu(data[0], dev), _gpu(data[1], dev)
        if inplace:
            plan.execute(re, im, batch=batch, inverse=inverse)
            return re.cpu().numpy() + 1j * im.cpu().numpy()
        ore, oim = torch.empty_like(re), torch.empty_like(im)
        plan.execute(re, im, ore, oim, batch=batch, inverse=inverse)
        assert np.array_equal(re.cpu().numpy(), data[0]) and np.array_equal(im.cpu().numpy(), data[1])
        return ore.cpu().numpy() + 1j * oim.cpu().numpy()
    a = _gpu(data, dev)
    if inplace:
        plan.execute(a, batch=batch, inverse=inverse)
        return a.cpu().numpy()
    b = torch.empty_like(a)
    plan.execute(a, b, batch=batch, inverse=inverse)
    assert np.array_equal(a.cpu().numpy(), data)
    return b.cpu().numpy()


# reference grid: 1D 2^{3,8,9,10,11,13}, 2D {2^4,2^7,2^8,2^10}^2, 3D {2^4,2^7}^3 (the 2^10 3D cases and
# 1D 2^20 need the multi-pass path and are covered in test_large.py when present), capped at 32 MB
SHAPES_1D = [(1 << k,) for k in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14)]     # 2^14: the longest single-pass complex64 row
SHAPES_2D = [(1 << a, 1 << b) for a in (4, 7, 8, 10) for b in (4, 7, 8, 10)]
SHAPES_3D = [(1 << a, 1 << b, 1 << c) for a in (4, 7) for b in (4, 7) for c in (4, 7)] + [(2, 2, 2), (4, 8, 2), (256, 4, 64)]
BATCHES = [1, 16, 128, 1024, 4096]      # test/test_errors.py:139


def _cases():
    out = []
    for shape in SHAPES_1D + SHAPES_2D + SHAPES_3D:
        size = int(np.prod(shape))
        for batch in BATCHES:
            if size * batch * 16 > 32 * 1024 * 1024:       # reference's default buffer budget (test_errors.py:142-145)
                continue
            out.append((shape, batch))
    return out


@pytest.mark.parametrize("shape,batch", _cases(), ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) else "b%s" % v)
@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.float64, np.complex128], ids=["f32split", "c64", "f64split", "c128"])
def test_grid_vs_numpy(cuda_device, shape, batch, dtype):
    """test/test_errors.py:18-114 re-expressed with seeded inputs and both tolerances."""
    from pyfft_b200.cuda import Plan
    size = int(np.prod(shape))
    data = no.make_input(shape, batch, dtype, seed=1000 + size % 97 + batch)
    z = (data[0] + 1j * data[1]) if isinstance(data, tuple) else data
    plan = Plan(shape, dtype=dtype, normalize=True, wait_for_finish=True)
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


RESTATE = [((1024,), 16), ((4096,), 8), ((8192,), 2), ((64, 256), 4), ((1024, 64), 1), ((16, 32, 64), 2), ((128, 16, 16), 1)]


@pytest.mark.parametrize("shape,batch", RESTATE)
@pytest.mark.parametrize("dtype", [np.complex64, np.float32, np.complex128])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("normalize,scale", [(True, 1.0), (False, 1.0), (True, 3.0)])
def test_vs_reference_restatement(cuda_device, shape, batch, dtype, inverse, normalize, scale):
    """CUDA path vs the restated reference algorithm on identical inputs, all scaling modes."""
    from pyfft_b200.cuda import Plan
    size = int(np.prod(shape))
    data = no.make_input(shape, batch, dtype, seed=31)
    if isinstance(data, tuple):
        re, im = data
    else:
        re, im = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
    pre, pim = pr.pyfft_execute(re, im, shape, batch, inverse, normalize, scale)
    plan = Plan(shape, dtype=dtype, normalize=normalize, scale=scale)
    got = _run(plan, cuda_device, data, batch, inverse, inplace=True)
    want64 = no.fft_oracle(re.astype(np.float64) + 1j * im.astype(np.float64), shape, batch, inverse, normalize, scale)
    tol = no.tolerance(dtype, size)
    assert no.rel_l2(got, pre + 1j * pim) < tol
    assert no.rel_l2(got, want64) < tol
    # we should be at least as close to the exact result as the restated reference is (x1.5 slack)
    assert no.rel_l2(got, want64) < 2.0 * no.rel_l2(pre + 1j * pim, want64) + 1e-16


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures(cuda_device, path):
    from pyfft_b200.cuda import Plan
    g = np.load(path)
    shape, batch = tuple(int(s) for s in g["shape"]), int(g["batch"])
    dtype = np.dtype(str(g["dtype"]))
    inverse, normalize, scale = bool(g["inverse"]), bool(g["normalize"]), float(g["scale"])
    if dtype.kind == "c":
        data = (g["re"] + 1j * g["im"]).astype(dtype)
    else:
        data = (g["re"], g["im"])
    plan = Plan(shape, dtype=dtype, normalize=normalize, scale=scale)
    got = _run(plan, cuda_device, data, batch, inverse, inplace=False)
    tol = no.tolerance(dtype, int(np.prod(shape)))
    assert no.rel_l2(got, g["expect64"]) < tol
    assert no.rel_l2(got, g["pyfft_re"] + 1j * g["pyfft_im"]) < tol


def test_batch_tail_and_changing_batch(cuda_device):
    """One plan serves any batch (plan.py:179-192); batches that do not fill the last CTA."""
    from pyfft_b200.cuda import Plan
    plan = Plan(256, dtype=np.complex64)
    for batch in (1, 3, 7, 8, 9, 31, 1000):
        data = no.make_input((256,), batch, np.complex64, seed=batch)
        got = _run(plan, cuda_device, data, batch, False, inplace=True)
        assert no.rel_l2(got, no.fft_oracle(data, (256,), batch)) < no.tolerance(np.complex64, 256)
    plan = Plan((32, 16), dtype=np.float64)
    for batch in (1, 5):
        data = no.make_input((32, 16), batch, np.float64, seed=batch)
        got = _run(plan, cuda_device, data, batch, False, inplace=False)
        assert no.rel_l2(got, no.fft_oracle(data[0] + 1j * data[1], (32, 16), batch)) < no.tolerance(np.float64, 512)


def test_guard_region_untouched(cuda_device):
    """Transforms write exactly their own elements: guard bands around the buffers stay intact."""
    import torch
    from pyfft_b200.cuda import Plan
    shape, batch = (64, 32), 3
    n = 64 * 32 * batch
    plan = Plan(shape, dtype=np.complex64)
    buf = torch.full((n + 512,), 7 + 7j, dtype=torch.complex64, device=cuda_device)
    data = no.make_input(shape, batch, np.complex64, seed=5)
    buf[256:256 + n] = _gpu(data.ravel(), cuda_device)
    view = buf[256:256 + n]
    plan.execute(view, batch=batch)
    out = buf.cpu().numpy()
    assert np.all(out[:256] == 7 + 7j) and np.all(out[256 + n:] == 7 + 7j)
    assert no.rel_l2(out[256:256 + n].reshape(data.shape), no.fft_oracle(data, shape, batch)) < 2e-4


def test_every_kernel_variant(cuda_device):
    """Each registered kernel variant (defaults, alternates, tuning candidates, TMA-staged ones) run as
    a single pass through b2fft_run_variant, forward / inverse / split, against numpy.fft."""
    import ctypes
    import torch
    from pyfft_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(256)
    stream = torch.cuda.current_stream().cuda_stream
    n_checked = 0
    for i in range(lib.b2fft_num_variants()):
        lib.b2fft_variant_info(i, buf, len(buf))
        f = buf.value.decode().split()
        name, prec, lg, W, G = f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4])
        n = 1 << lg
        inner = 1 if W == 1 else 2 * W
        outer = 2 * G + 1 if W == 1 else 3          # odd tile count: tail CTA / partial TMA group
        cdt = np.complex64 if prec == 0 else np.complex128
        rng = np.random.default_rng(i)
        x = (rng.standard_normal((outer, n, inner)) + 1j * rng.standard_normal((outer, n, inner))).astype(cdt)
        want = np.fft.fft(x.astype(np.complex128), axis=1)
        tol = no.tolerance(cdt, n)
        n_tiles = outer * (inner // W)
        a = _gpu(x, cuda_device)
        for inverse in (0, 1):
            b = torch.zeros_like(a)
            _lib.check(lib.b2fft_run_variant(i, a.data_ptr(), None, b.data_ptr(), None, 0, inverse, n_tiles, inner, 0, stream))
            ref = np.fft.ifft(x.astype(np.complex128), axis=1) * n if inverse else want
            assert no.rel_l2(b.cpu().numpy(), ref) < tol, (name, inverse)
        re, im = _gpu(x.real.copy(), cuda_device), _gpu(x.imag.copy(), cuda_device)
        ore, oim = torch.zeros_like(re), torch.zeros_like(im)
        rc = lib.b2fft_run_variant(i, re.data_ptr(), im.data_ptr(), ore.data_ptr(), oim.data_ptr(), 1, 0, n_tiles, inner, 0, stream)
        if "_fused2" in name:                 # fused two-step kernels are interleaved-only; they must say so
            assert rc != _lib.OK and "not supported" in _lib.last_error()
            lib.b2fft_run_variant(i, a.data_ptr(), None, a.data_ptr(), None, 0, 0, n_tiles, inner, 0, stream)   # in place
            assert no.rel_l2(a.cpu().numpy(), want) < tol, (name, "in place")
        else:
            _lib.check(rc)
            assert no.rel_l2(ore.cpu().numpy() + 1j * oim.cpu().numpy(), want) < tol, (name, "split")
        n_checked += 1
    assert n_checked >= 60


@pytest.mark.parametrize("chunk_mb", [0, 1, 32])
def test_l2_chunked_schedule_is_exact(cuda_device, chunk_mb):
    """The L2-resident chunked schedule only reorders independent work: results are bit-identical
    for every chunk size (0 = whole-array passes), in-place and out-of-place, 2D and 3D."""
    from pyfft_b200 import _lib
    from pyfft_b200.cuda import Plan
    lib = _lib.load()
    results = []
    for mb in (0, chunk_mb):
        _lib.check(lib.b2fft_set_option(b"l2_chunk_bytes", float(mb) * 1024 * 1024))
        outs = []
        for shape, batch, dtype in (((256, 512), 5, np.complex64), ((32, 64, 128), 3, np.complex64),
                                    ((64, 64, 64), 2, np.complex128), ((128, 256), 7, np.float32)):
            data = no.make_input(shape, batch, dtype, seed=11)
            plan = Plan(shape, dtype=dtype)
            outs.append(_run(plan, cuda_device, data, batch, False, inplace=True))
            outs.append(_run(plan, cuda_device, data, batch, True, inplace=False))
        results.append(outs)
    _lib.check(lib.b2fft_set_option(b"l2_chunk_bytes", 32.0 * 1024 * 1024))
    for a, b in zip(*results):
        assert np.array_equal(a, b)
    z = no.make_input((256, 512), 5, np.complex64, seed=11)
    assert no.rel_l2(results[1][0], no.fft_oracle(z, (256, 512), 5)) < no.tolerance(np.complex64, 256 * 512)


def test_persistent_tma_kernels_are_race_free(cuda_device):
    """Race detector for the persistent TMA-staged kernels (row and strided): many launches over many
    groups per CTA must be bit-identical run to run and match numpy.  (A missing generic->async proxy
    fence before the slot refill showed up here as sporadically corrupted rows; see kernels.cuh.)"""
    import ctypes
    import torch
    from pyfft_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(256)
    stream = torch.cuda.current_stream().cuda_stream
    checked = 0
    for i in range(lib.b2fft_num_variants()):
        lib.b2fft_variant_info(i, buf, len(buf))
        f = buf.value.decode().split()
        name, prec, lg, W, G = f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4])
        if "_tma" not in name or lg < 8:
            continue
        n = 1 << lg
        inner = 1 if W == 1 else 512
        outer = 4096 if W == 1 else max(2, 4096 * 1024 // (n * inner))
        cdt = np.complex64 if prec == 0 else np.complex128
        rng = np.random.default_rng(i)
        x = (rng.standard_normal((outer, n, inner)) + 1j * rng.standard_normal((outer, n, inner))).astype(cdt)
        a = _gpu(x, cuda_device)
        n_tiles = outer * (inner // W)
        first = None
        for rep in range(12):
            inplace = rep % 2 == 1
            b = a.clone() if inplace else torch.zeros_like(a)
            src = b if inplace else a
            _lib.check(lib.b2fft_run_variant(i, src.data_ptr(), None, b.data_ptr(), None, 0, 0, n_tiles, inner, 0, stream))
            if first is None:
                first = b
            else:
                assert torch.equal(first, b), (name, rep)
        want = np.fft.fft(x.astype(np.complex128), axis=1)
        assert no.rel_l2(first.cpu().numpy(), want) < no.tolerance(cdt, n), name
        checked += 1
    assert checked >= 12
