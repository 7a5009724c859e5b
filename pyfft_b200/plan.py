"""FFTPlan: the host-side mirror of the reference's ``pyfft.plan.FFTPlan``.

Same constructor arguments, the same two ``execute`` signatures, the same exception types
and scaling rules (reference pyfft/plan.py:66-284, pyfft/kernel.py:23-37); the work itself
is one ``b2fft_execute`` call into libb2fft.so (include/b2fft.h).
"""
import ctypes

import numpy

from . import _lib

_NP_DTYPES = {
    numpy.dtype(numpy.complex64): (_lib.F32, _lib.INTERLEAVED),
    numpy.dtype(numpy.float32): (_lib.F32, _lib.SPLIT),
    numpy.dtype(numpy.complex128): (_lib.F64, _lib.INTERLEAVED),
    numpy.dtype(numpy.float64): (_lib.F64, _lib.SPLIT),
}


def _resolve_dtype(dtype):
    """Accept numpy dtypes/types, strings and torch dtypes; anything else is a ValueError
    (reference pyfft/plan.py:26-48)."""
    try:
        import torch
        if isinstance(dtype, torch.dtype):
            dtype = {torch.complex64: numpy.complex64, torch.float32: numpy.float32,
                     torch.complex128: numpy.complex128, torch.float64: numpy.float64}.get(dtype, dtype)
            if isinstance(dtype, torch.dtype):
                raise ValueError("Data type " + str(dtype) + " is not supported")
    except ImportError:
        pass
    try:
        dt = numpy.dtype(dtype)
    except TypeError:
        raise ValueError("Data type " + str(dtype) + " is not supported")
    if dt not in _NP_DTYPES:
        raise ValueError("Data type " + str(dtype) + " is not supported")
    return dt


def _normalize_shape(shape):
    """int | 1/2/3-tuple in numpy order -> (x, y, z), x = last (contiguous) axis
    (reference pyfft/plan.py:73-89)."""
    if isinstance(shape, bool):
        raise ValueError("Wrong shape")
    if isinstance(shape, (int, numpy.integer)):
        return 1, (int(shape), 1, 1)
    if isinstance(shape, (tuple, list)):
        dims = []
        for s in shape:
            if isinstance(s, bool) or not isinstance(s, (int, numpy.integer)):
                raise ValueError("Wrong shape")
            dims.append(int(s))
        if len(dims) == 1:
            return 1, (dims[0], 1, 1)
        if len(dims) == 2:
            return 2, (dims[1], dims[0], 1)
        if len(dims) == 3:
            return 3, (dims[2], dims[1], dims[0])
    raise ValueError("Wrong shape")


def _device_pointer(obj, what):
    """Raw device address of a buffer argument.

    Accepts torch CUDA tensors, anything with ``__cuda_array_interface__`` (CuPy, Numba),
    PyCUDA GPUArray / DeviceAllocation (``.gpudata`` / ``int()``), objects with ``data_ptr()``
    or ``.ptr``, and plain ints.  (The reference unwraps GPUArray.gpudata, pyfft/cuda.py:35-39.)
    """
    if obj is None:
        return None, None, None
    if isinstance(obj, (int, numpy.integer)) and not isinstance(obj, bool):
        return int(obj), None, None
    try:
        import torch
        if isinstance(obj, torch.Tensor):
            if not obj.is_cuda:
                raise ValueError("%s: expected a CUDA tensor, got a %s tensor" % (what, obj.device.type))
            if not obj.is_contiguous():
                raise ValueError("%s: tensor must be contiguous" % what)
            return obj.data_ptr(), obj.numel() * obj.element_size(), obj.dtype
    except ImportError:
        pass
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        nbytes, npdt = None, None
        try:
            npdt = numpy.dtype(cai["typestr"])
            nbytes = int(numpy.prod(cai["shape"])) * npdt.itemsize
        except Exception:
            pass
        strides = cai.get("strides")
        if strides is not None and npdt is not None:          # None means C-contiguous; anything else must equal it
            expect, acc = [], npdt.itemsize
            for dim in reversed(tuple(cai["shape"])):
                expect.append(acc)
                acc *= int(dim)
            dense = all(int(d) <= 1 or int(st) == e for d, st, e in zip(reversed(tuple(cai["shape"])), reversed(tuple(strides)), expect))
            if not dense:
                raise ValueError("%s: array must be C-contiguous (strides %r for shape %r)" % (what, strides, cai["shape"]))
        return int(cai["data"][0]), nbytes, npdt
    if hasattr(obj, "gpudata"):                              # pycuda.gpuarray.GPUArray (pyfft/cuda.py:37-39)
        flags = getattr(obj, "flags", None)
        if flags is not None and not getattr(flags, "c_contiguous", True):
            raise ValueError("%s: GPUArray must be C-contiguous" % what)
        return int(obj.gpudata), getattr(obj, "nbytes", None), getattr(obj, "dtype", None)
    if hasattr(obj, "data_ptr"):
        return int(obj.data_ptr()), None, None
    if hasattr(obj, "ptr"):
        return int(obj.ptr), None, None
    try:
        return int(obj), None, None
    except Exception:
        raise TypeError("%s: cannot obtain a device pointer from %r" % (what, type(obj)))


def _stream_handle(stream):
    """cudaStream_t value of a stream-like object (torch / PyCUDA / CuPy / int)."""
    if stream is None:
        return 0
    if isinstance(stream, (int, numpy.integer)):
        return int(stream)
    for attr in ("cuda_stream", "handle", "ptr"):
        if hasattr(stream, attr):
            return int(getattr(stream, attr))
    raise TypeError("cannot obtain a cudaStream_t from %r" % type(stream))


class FFTPlan(object):
    """Plan preparation and execution (reference pyfft/plan.py:66-284)."""

    def __init__(self, context, shape, dtype=numpy.complex64, normalize=True,
                 wait_for_finish=None, fast_math=True, scale=1.0):
        self._dim, (x, y, z) = _normalize_shape(shape)
        self._xyz = (x, y, z)
        size = x * y * z
        if x < 1 or y < 1 or z < 1 or (size & (size - 1)) != 0:
            raise ValueError("Array dimensions must be powers of two")      # plan.py:23-24
        self._dtype = _resolve_dtype(dtype)
        self._prec, self._layout = _NP_DTYPES[self._dtype]
        self._split = self._layout == _lib.SPLIT
        self._size = size
        self._context = context
        self._normalize = bool(normalize)
        self._scale = float(scale)
        self._fast_math = bool(fast_math)
        self._wait_for_finish = wait_for_finish
        self._itemsize = self._dtype.itemsize
        self._workspace = None
        self._workspace_bytes = 0
        self._last_batch_size = None

        self._lib = _lib.load()                # raises if the CUDA library is not built
        self._device = context.device_index()
        handle = ctypes.c_void_p()
        dims = (ctypes.c_int64 * 3)(x, y, z)
        _lib.check(self._lib.b2fft_plan_create(ctypes.byref(handle), self._dim, dims, self._prec, self._layout,
                                               int(self._normalize), self._scale, int(self._fast_math),
                                               self._device))
        self._handle = handle

        # execute() signature depends on the layout (plan.py:104-107)
        if self._split:
            self.execute = self._executeSplit
        else:
            self.execute = self._executeInterleaved

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self._lib.b2fft_plan_destroy(h)
            except Exception:
                pass
            self._handle = None

    # ------------------------------------------------------------------ introspection
    @property
    def passes(self):
        buf = ctypes.create_string_buffer(4096)
        _lib.check(self._lib.b2fft_plan_describe(self._handle, buf, len(buf)))
        return [l for l in buf.value.decode().splitlines() if l]

    @property
    def launch_count(self):
        return int(self._lib.b2fft_plan_launch_count(self._handle))

    # ------------------------------------------------------------------ execution
    def _check_buffer(self, obj, what, batch):
        ptr, nbytes, tdtype = _device_pointer(obj, what)
        need = self._size * batch * self._itemsize
        if nbytes is not None and nbytes < need:
            raise ValueError("%s holds %d bytes, the plan needs %d (shape %s, batch %d)"
                             % (what, nbytes, need, self._xyz, batch))
        if tdtype is not None:
            import torch
            if isinstance(tdtype, torch.dtype):
                want = {numpy.dtype(numpy.complex64): torch.complex64, numpy.dtype(numpy.float32): torch.float32,
                        numpy.dtype(numpy.complex128): torch.complex128, numpy.dtype(numpy.float64): torch.float64}
                if tdtype != want[self._dtype]:
                    raise TypeError("%s has dtype %s, the plan was created for %s" % (what, tdtype, self._dtype))
            else:                                    # numpy dtype of a CAI / GPUArray buffer; raw byte buffers are accepted
                try:
                    npdt = numpy.dtype(tdtype)
                except TypeError:
                    npdt = None
                if npdt is not None and npdt.kind in "fc" and npdt != self._dtype:
                    raise TypeError("%s has dtype %s, the plan was created for %s" % (what, npdt, self._dtype))
        return ptr

    def _ensure_workspace(self, batch, in_place):
        """Workspace is owned by the caller side of the C ABI (plan.py:184-192 re-allocates the
        reference's temp buffer per batch size the same way).  Only plans with an axis too long
        for one pass need one, and out-of-place executes of a long first axis do not."""
        key = (batch, bool(in_place))
        if key == self._last_batch_size:
            return
        need = ctypes.c_size_t(0)
        _lib.check(self._lib.b2fft_plan_workspace_bytes_ex(self._handle, batch, int(bool(in_place)),
                                                           ctypes.byref(need)))
        if need.value > self._workspace_bytes:
            self._workspace = None                      # release the old buffer before asking for a bigger one
            self._workspace = self._context.allocate(need.value)
            ptr, _, _ = _device_pointer(self._workspace, "workspace")
            _lib.check(self._lib.b2fft_plan_set_workspace(self._handle, ptr, need.value))
            self._workspace_bytes = need.value
        self._last_batch_size = key

    def _execute(self, wait_for_finish, inverse, batch, in0, in1, out0, out1):
        batch = int(batch)
        if batch < 0:
            raise ValueError("batch must be non-negative")
        p_in0 = self._check_buffer(in0, "data_in" + ("_re" if self._split else ""), batch)
        p_out0 = self._check_buffer(out0, "data_out" + ("_re" if self._split else ""), batch)
        p_in1 = p_out1 = None
        if self._split:
            p_in1 = self._check_buffer(in1, "data_in_im", batch)
            p_out1 = self._check_buffer(out1, "data_out_im", batch)
        in_place = p_in0 == p_out0 or (self._split and p_in1 == p_out1)
        self._ensure_workspace(batch, in_place)
        stream = self._context.get_stream()
        ws = self._workspace
        if ws is not None and hasattr(ws, "record_stream"):
            try:                                   # torch workspace used on a torch stream other than the allocating one
                import torch
                if isinstance(stream, torch.cuda.Stream):
                    ws.record_stream(stream)
            except ImportError:
                pass
        _lib.check(self._lib.b2fft_execute(self._handle, p_in0, p_in1, p_out0, p_out1, int(bool(inverse)), batch,
                                           _stream_handle(stream)))
        # the execute kwarg wins over the constructor's setting (plan.py:250-253)
        wait = self._wait_for_finish
        if wait_for_finish is not None:
            wait = wait_for_finish
        if wait:
            self._context.wait(stream)
            return None
        return stream                                                    # plan.py:255-259

    def _executeInterleaved(self, data_in, data_out=None, inverse=False, batch=1, wait_for_finish=None):
        """Execute plan for interleaved complex arrays (plan.py:261-271)."""
        if data_out is None:
            data_out = data_in
        return self._execute(wait_for_finish, inverse, batch, data_in, None, data_out, None)

    def _executeSplit(self, data_in_re, data_in_im, data_out_re=None, data_out_im=None, inverse=False, batch=1,
                      wait_for_finish=None):
        """Execute plan for split re/im arrays (plan.py:273-284)."""
        if data_out_re is None and data_out_im is None:
            data_out_re = data_in_re
            data_out_im = data_in_im
        elif data_out_re is None or data_out_im is None:
            raise ValueError("both output arrays (or neither) must be given for a split plan")
        return self._execute(wait_for_finish, inverse, batch, data_in_re, data_in_im, data_out_re, data_out_im)
