"""GPU tests: the reference's API-contract suite (test/test_functionality.py) re-expressed for
pyfft_b200.cuda.Plan with torch tensors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PREC = [(np.float32, np.complex64), (np.float64, np.complex128)]


def _gpu(arr, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).to(dev)


@pytest.mark.parametrize("scalar,cplx", PREC)
def test_shapes_and_types(cuda_device, scalar, cplx):
    """test_functionality.py:11-22"""
    from pyfft_b200.cuda import Plan
    for shape in [16, (16,), (16, 16), (16, 16, 16)]:
        Plan(shape, dtype=scalar, context=0)
    for dtype in (scalar, cplx):
        Plan((16, 16), dtype=dtype, context=0)


@pytest.mark.parametrize("scalar,cplx", PREC)
def test_execute_signature_split(cuda_device, scalar, cplx):
    """test_functionality.py:24-36"""
    from pyfft_b200.cuda import Plan
    plan = Plan((16,), dtype=scalar, context=0)
    a, b, c, d = (_gpu(np.ones(16, dtype=scalar), cuda_device) for _ in range(4))
    plan.execute(a, b)
    plan.execute(a, b, c, d)
    plan.execute(a, b, a, b)


@pytest.mark.parametrize("scalar,cplx", PREC)
def test_execute_signature_interleaved(cuda_device, scalar, cplx):
    """test_functionality.py:38-51"""
    from pyfft_b200.cuda import Plan
    plan = Plan((16,), dtype=cplx, context=0)
    a, b = (_gpu(np.ones(16, dtype=cplx), cuda_device) for _ in range(2))
    plan.execute(a)
    plan.execute(a, b)
    plan.execute(a, a)
    with pytest.raises(TypeError):
        plan.execute(a, b, a, b, inverse=True)


@pytest.mark.parametrize("scalar,cplx", PREC)
def test_normalize(cuda_device, scalar, cplx):
    """test_functionality.py:53-77 (known answer: ones(16))"""
    from pyfft_b200.cuda import Plan
    data = np.ones(16, dtype=cplx)
    for normalize in (True, False):
        plan = Plan(data.shape, normalize=normalize, dtype=cplx, context=0)
        a = _gpu(data, cuda_device)
        plan.execute(a)
        res = a.cpu().numpy()
        assert np.sum(np.abs(np.fft.fft(data) - res)) / data.size < 1e-6
        plan.execute(a, inverse=True)
        res = a.cpu().numpy()
        coeff = 1 if normalize else data.size
        assert np.sum(np.abs(data * coeff - res)) / data.size < 1e-6


@pytest.mark.parametrize("scalar,cplx", PREC)
@pytest.mark.parametrize("fast_math", [True, False])
def test_scale(cuda_device, scalar, cplx, fast_math):
    """test_functionality.py:79-100"""
    from pyfft_b200.cuda import Plan
    data = np.ones(16, dtype=cplx)
    for scale in (1.0, 10.0):
        plan = Plan(data.shape, scale=scale, dtype=cplx, context=0, normalize=True, fast_math=fast_math)
        a = _gpu(data, cuda_device)
        plan.execute(a)
        assert np.sum(np.abs(np.fft.fft(data) * scale - a.cpu().numpy())) / data.size < 1e-6
        plan.execute(a, inverse=True)
        assert np.sum(np.abs(data - a.cpu().numpy())) / data.size < 1e-6


@pytest.mark.parametrize("scalar,cplx", PREC)
def test_fast_math(cuda_device, scalar, cplx):
    """test_functionality.py:102-115"""
    from pyfft_b200.cuda import Plan
    data = np.ones(8192, dtype=cplx)
    for fast_math in (True, False):
        plan = Plan(data.shape, normalize=True, dtype=cplx, context=0, fast_math=fast_math)
        a = _gpu(data, cuda_device)
        plan.execute(a)
        plan.execute(a, inverse=True)
        assert np.sum(np.abs(data - a.cpu().numpy())) / data.size < 1e-6


@pytest.mark.parametrize("scalar,cplx", PREC)
def test_allocation_and_doc_example(cuda_device, scalar, cplx):
    """test_functionality.py:117-127, doc/source/index.rst:61-99"""
    from pyfft_b200.cuda import Plan
    plan = Plan((32, 32, 32), dtype=cplx, context=0)
    a = _gpu(np.ones((32, 32, 32), dtype=cplx), cuda_device)
    plan.execute(a)
    res = a.cpu().numpy()
    assert abs(res[0, 0, 0] - 32 ** 3) < 1e-2 and np.abs(res).sum() - abs(res[0, 0, 0]) < 1e-2
    plan = Plan((16, 16), dtype=cplx)
    data = np.ones((16, 16), dtype=cplx)
    g = _gpu(data, cuda_device)
    plan.execute(g)
    assert abs(g.cpu().numpy()[0, 0] - 256) < 1e-4
    plan.execute(g, inverse=True)
    assert np.abs(g.cpu().numpy() - data).sum() / data.size < 1e-6


def test_wrong_arguments(cuda_device):
    """test_functionality.py:129-139"""
    from pyfft_b200.cuda import Plan
    with pytest.raises(ValueError):
        Plan((17,), dtype=np.complex64)
    with pytest.raises(ValueError):
        Plan((16,), dtype=np.int32)
    with pytest.raises(ValueError):
        Plan((16, 16, 16, 16), dtype=np.complex64)
    with pytest.raises(ValueError):
        Plan("16", dtype=np.complex64)


class _Pool(object):
    """stands in for pycuda.tools.DeviceMemoryPool (test_functionality.py:147-150)"""

    def __init__(self):
        self.calls = 0

    def allocate(self, nbytes):
        import torch
        self.calls += 1
        return torch.empty(nbytes, dtype=torch.uint8, device="cuda")


def test_mempool_stream_and_wait(cuda_device):
    """test_functionality.py:147-165"""
    import torch
    from pyfft_b200.cuda import Plan
    Plan((32, 32, 32), dtype=np.complex64, mempool=_Pool())
    stream = torch.cuda.Stream()
    plan = Plan((32, 32, 32), dtype=np.complex64, stream=stream)
    a = _gpu(np.ones((32, 32, 32), dtype=np.complex64), cuda_device)
    stream.wait_stream(torch.cuda.current_stream())
    ret = plan.execute(a)
    assert ret is stream                      # external stream => wait_for_finish defaults to False
    stream.synchronize()
    assert abs(a.cpu().numpy()[0, 0, 0] - 32 ** 3) < 1e-2
    plan = Plan((32, 32, 32), dtype=np.complex64)
    a = _gpu(np.ones((32, 32, 32), dtype=np.complex64), cuda_device)
    assert plan.execute(a) is None            # default: waits
    s = plan.execute(a, wait_for_finish=False)
    s.synchronize()
    # raw cudaStream_t handle
    plan = Plan((16,), dtype=np.complex64, stream=int(stream.cuda_stream), wait_for_finish=True)
    b = _gpu(np.ones(16, dtype=np.complex64), cuda_device)
    torch.cuda.synchronize()
    assert plan.execute(b) is None
    assert abs(b.cpu().numpy()[0] - 16) < 1e-5


def test_raw_pointers_and_cuda_array_interface(cuda_device):
    from pyfft_b200.cuda import Plan
    data = (np.arange(64) + 1j * np.arange(64)[::-1]).astype(np.complex64)
    a = _gpu(data, cuda_device)
    b = _gpu(np.zeros(64, np.complex64), cuda_device)
    plan = Plan(64, dtype=np.complex64)
    plan.execute(int(a.data_ptr()), int(b.data_ptr()))
    assert np.allclose(b.cpu().numpy(), np.fft.fft(data), rtol=1e-5, atol=1e-3)
    assert np.array_equal(a.cpu().numpy(), data)      # out-of-place leaves the input intact

    class CAI(object):
        def __init__(self, t):
            self.__cuda_array_interface__ = t.__cuda_array_interface__
    c = _gpu(np.zeros(64, np.complex64), cuda_device)
    plan.execute(CAI(a), CAI(c))
    assert np.allclose(c.cpu().numpy(), np.fft.fft(data), rtol=1e-5, atol=1e-3)


def test_buffer_checks(cuda_device):
    from pyfft_b200.cuda import Plan
    plan = Plan(64, dtype=np.complex64)
    small = _gpu(np.zeros(32, np.complex64), cuda_device)
    with pytest.raises(ValueError):
        plan.execute(small)
    wrong = _gpu(np.zeros(64, np.complex128), cuda_device)
    with pytest.raises(TypeError):
        plan.execute(wrong)
    import torch
    with pytest.raises(ValueError):
        plan.execute(torch.zeros(64, dtype=torch.complex64))      # CPU tensor: no fallback
    ok = _gpu(np.zeros(64 * 3, np.complex64), cuda_device)
    with pytest.raises(ValueError):
        plan.execute(ok, batch=4)
    plan.execute(ok, batch=3)


def test_native_library_is_loaded(cuda_device):
    """The GPU tests must exercise the in-tree CUDA library, not a fallback."""
    from pyfft_b200 import _lib
    from pyfft_b200.cuda import Plan
    plan = Plan(1024, dtype=np.complex64)
    import torch
    a = torch.zeros(1024, dtype=torch.complex64, device=cuda_device)
    before = plan.launch_count
    plan.execute(a)
    assert plan.launch_count == before + 1
    maps = open("/proc/self/maps").read()
    assert _lib.LIB_PATH in maps
