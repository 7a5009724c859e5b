"""CPU tests of the Python host layer that mirrors the reference's interface (no GPU, no kernels):
shape / dtype normalisation (pyfft/plan.py:23-48,73-89), buffer and stream unwrapping
(pyfft/cuda.py:35-46: GPUArray.gpudata; here also torch, __cuda_array_interface__, raw pointers)."""
import numpy as np
import pytest
import torch

from pyfft_b200.plan import _device_pointer, _normalize_shape, _resolve_dtype, _stream_handle


def test_shape_normalisation_matches_reference():
    # plan.py:73-89: int | (x,) | (y, x) | (z, y, x), x = last (contiguous) axis
    assert _normalize_shape(16) == (1, (16, 1, 1))
    assert _normalize_shape((16,)) == (1, (16, 1, 1))
    assert _normalize_shape((8, 16)) == (2, (16, 8, 1))
    assert _normalize_shape([4, 8, 16]) == (3, (16, 8, 4))
    assert _normalize_shape(np.int64(32)) == (1, (32, 1, 1))
    for bad in ("16", (1, 2, 3, 4), (), None, 1.5, (16, "8"), True, (True, 4)):
        with pytest.raises(ValueError):
            _normalize_shape(bad)


def test_dtype_resolution():
    # plan.py:26-48: complex dtypes = interleaved, real dtypes = split; everything else is a ValueError
    for ok in (np.complex64, np.complex128, np.float32, np.float64, "complex64", np.dtype("float64"),
               torch.complex64, torch.float32, torch.complex128, torch.float64):
        assert _resolve_dtype(ok).kind in "cf"
    for bad in (np.int32, np.float16, "int8", torch.int64, torch.bfloat16, object):
        with pytest.raises(ValueError):
            _resolve_dtype(bad)


class _GpuArray(object):            # pycuda.gpuarray.GPUArray look-alike (cuda.py:37-39 unwraps .gpudata)
    def __init__(self, ptr, nbytes):
        self.gpudata, self.nbytes = ptr, nbytes


class _Cai(object):                 # CuPy / Numba look-alike
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<c8", "data": (ptr, False), "version": 2}


class _Ptr(object):
    def __init__(self, p):
        self.ptr = p


def test_device_pointer_unwrapping():
    assert _device_pointer(None, "x") == (None, None, None)
    assert _device_pointer(0x7f0000001000, "x")[0] == 0x7f0000001000
    assert _device_pointer(np.int64(4096), "x")[0] == 4096
    assert _device_pointer(_GpuArray(8192, 64), "x")[:2] == (8192, 64)
    assert _device_pointer(_Cai(12288, 10), "x")[:2] == (12288, 80)
    assert _device_pointer(_Ptr(16384), "x")[0] == 16384
    with pytest.raises(ValueError):                       # host tensors are refused, there is no CPU path
        _device_pointer(torch.zeros(4, dtype=torch.complex64), "data_in")
    with pytest.raises(TypeError):
        _device_pointer(object(), "data_in")


class _PyCudaStream(object):
    handle = 1234


class _TorchLikeStream(object):
    cuda_stream = 5678


def test_stream_handles():
    assert _stream_handle(None) == 0
    assert _stream_handle(42) == 42
    assert _stream_handle(_PyCudaStream()) == 1234
    assert _stream_handle(_TorchLikeStream()) == 5678
    assert _stream_handle(_Ptr(99)) == 99
    with pytest.raises(TypeError):
        _stream_handle(object())


class _CaiView(object):             # non-contiguous CuPy-style view: strides given explicitly
    def __init__(self, ptr, shape, strides, typestr="<c8"):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": strides}


class _Flags(object):
    def __init__(self, c):
        self.c_contiguous = c


def test_noncontiguous_buffers_are_refused():
    # dense strides written out explicitly are fine, views with gaps are not (they would be transformed as if dense)
    assert _device_pointer(_CaiView(4096, (4, 8), (64, 8)), "x")[:2] == (4096, 256)
    assert _device_pointer(_CaiView(4096, (1, 8), (640, 8)), "x")[0] == 4096           # size-1 axes: any stride
    with pytest.raises(ValueError):
        _device_pointer(_CaiView(4096, (4, 8), (128, 8)), "x")
    with pytest.raises(ValueError):
        _device_pointer(_CaiView(4096, (4, 8), (64, 16)), "x")
    ga = _GpuArray(8192, 64)
    ga.flags = _Flags(False)
    with pytest.raises(ValueError):
        _device_pointer(ga, "x")
    ga.flags = _Flags(True)
    assert _device_pointer(ga, "x")[0] == 8192


def test_context_selector_must_be_understood():
    """`context=` objects that do not identify a device raise instead of silently meaning 'current device'."""
    from pyfft_b200 import cuda as b2cuda

    class _Dev(object):
        pass

    class _Ctx(object):
        def get_device(self):
            return _Dev()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            b2cuda._device_of(_Ctx(), None)
        return
    with pytest.raises(ValueError):
        b2cuda._device_of(_Ctx(), None)
