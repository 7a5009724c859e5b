"""``Plan`` factory, drop-in for ``pyfft.cuda.Plan`` (reference pyfft/cuda.py:116-138).

    Plan(shape, dtype=numpy.complex64, mempool=None, context=None, normalize=True,
         wait_for_finish=None, fast_math=True, stream=None, scale=1.0)

``stream``   torch.cuda.Stream, PyCUDA / CuPy stream or raw ``cudaStream_t`` int.  Given a
             stream, ``wait_for_finish`` defaults to False and ``execute`` returns the stream.
``context``  a device selector (int ordinal, ``torch.device``, or an object with
             ``get_device()``); ``wait_for_finish`` defaults to True.  As in the reference
             (pyfft/cuda.py:68-69,94-101) a plan created this way runs every ``execute`` on a fresh
             stream of that device (here it is first made to wait for the device's current torch
             stream, so the transform stays ordered after the kernels that produced its input).
neither      the plan runs on torch's *current* stream of the current device and waits.  (The
             reference creates a private stream here; running on the current stream instead keeps
             the transform ordered after the torch kernels that produced its input.)
``mempool``  object with ``allocate(nbytes)`` used for plan workspace instead of torch's
             caching allocator (reference pyfft/cuda.py:85-89).
"""
import numpy

from . import _lib
from .plan import FFTPlan


class Context(object):
    """Execution context handed to FFTPlan (reference pyfft/cuda.py:64-113)."""

    def __init__(self, device, stream, mempool, recreate_stream=False):
        self._device = device
        self._stream = stream
        self._mempool = mempool
        self._recreate_stream = bool(recreate_stream) and stream is None   # pyfft/cuda.py:68-69

    def device_index(self):
        return self._device

    def isCuda(self):
        return True

    def allocate(self, nbytes):
        if self._mempool is not None:
            return self._mempool.allocate(nbytes)
        import torch
        return torch.empty(int(nbytes), dtype=torch.uint8, device="cuda:%d" % self._device)

    def get_stream(self):
        if self._stream is not None:
            return self._stream
        import torch
        if self._recreate_stream:                        # pyfft/cuda.py:94-96 createQueue()
            s = torch.cuda.Stream(device=self._device)
            s.wait_stream(torch.cuda.current_stream(self._device))
            return s
        return torch.cuda.current_stream(self._device)

    def wait(self, stream):
        if hasattr(stream, "synchronize"):
            stream.synchronize()
        else:
            from .plan import _stream_handle
            _lib.check(_lib.load().b2fft_stream_synchronize(_stream_handle(stream)))


def _device_of(context_obj, stream_obj):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("pyfft_b200 needs a CUDA device; there is no CPU fallback")
    if context_obj is not None:
        if isinstance(context_obj, (int, numpy.integer)):
            return int(context_obj)
        if isinstance(context_obj, torch.device):
            return context_obj.index if context_obj.index is not None else torch.cuda.current_device()
        if isinstance(context_obj, str):
            d = torch.device(context_obj)
            return d.index if d.index is not None else torch.cuda.current_device()
        if hasattr(context_obj, "get_device"):            # pycuda.driver.Context -> pycuda.driver.Device
            dev = context_obj.get_device()
            if isinstance(dev, (int, numpy.integer)):
                return int(dev)
            for attr in ("index", "id"):
                if hasattr(dev, attr):
                    return int(getattr(dev, attr))
            if hasattr(dev, "pci_bus_id"):                # PyCUDA devices carry no ordinal: match the PCI bus id
                want = dev.pci_bus_id().lower().split(":", 1)[-1] if callable(dev.pci_bus_id) else str(dev.pci_bus_id).lower()
                for i in range(torch.cuda.device_count()):
                    p = torch.cuda.get_device_properties(i)
                    got = "%02x:%02x" % (getattr(p, "pci_bus_id", -1), getattr(p, "pci_device_id", -1))
                    if want.startswith(got) or got in want:
                        return i
        raise ValueError("context: cannot tell which CUDA device %r selects" % (context_obj,))
    if stream_obj is not None and hasattr(stream_obj, "device_index"):
        return int(stream_obj.device_index)
    if stream_obj is not None and hasattr(stream_obj, "device") and hasattr(stream_obj.device, "index") \
            and stream_obj.device.index is not None:
        return int(stream_obj.device.index)
    return torch.cuda.current_device()


def Plan(*args, **kwds):
    """Create an FFT plan; see the module docstring (reference pyfft/cuda.py:116-138)."""
    mempool = kwds.pop("mempool", None)
    context_obj = kwds.pop("context", None)
    stream_obj = kwds.pop("stream", None)

    if stream_obj is not None:
        wait_for_finish = False
    else:
        wait_for_finish = True

    if "wait_for_finish" not in kwds or kwds["wait_for_finish"] is None:
        kwds["wait_for_finish"] = wait_for_finish

    # argument errors (ValueError / TypeError) must surface before any CUDA work, as in the
    # reference where _FFTParams raises first (pyfft/plan.py:23-24,48,87-89)
    _validate_only(*args, **kwds)
    device = _device_of(context_obj, stream_obj)
    context = Context(device, stream_obj, mempool, recreate_stream=context_obj is not None)
    return FFTPlan(context, *args, **kwds)


def _validate_only(shape, dtype=numpy.complex64, normalize=True, wait_for_finish=None, fast_math=True, scale=1.0):
    from .plan import _normalize_shape, _resolve_dtype
    _, (x, y, z) = _normalize_shape(shape)
    size = x * y * z
    if x < 1 or y < 1 or z < 1 or (size & (size - 1)) != 0:
        raise ValueError("Array dimensions must be powers of two")
    _resolve_dtype(dtype)
