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
