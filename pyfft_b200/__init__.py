"""pyfft_b200 -- B200-native batched C2C FFT with pyfft's Plan/execute API.

Drop-in for ``pyfft.cuda.Plan`` (reference pyfft/cuda.py:116-138):

    from pyfft_b200.cuda import Plan
    plan = Plan((1024, 1024), dtype=numpy.complex64, stream=stream)
    plan.execute(gpu_data)                       # in place, forward
    plan.execute(gpu_data, gpu_out, inverse=True)

The compute path is hand-written sm_100a CUDA in ``libb2fft.so`` (pyfft_b200/csrc), reached
through the C ABI in include/b2fft.h.  There is no CPU fallback: importing the package
works anywhere, but creating a plan without the built library or without a GPU raises.
"""
VERSION = (0, 1, 0)          # reference: pyfft/__init__.py:1 VERSION = (0, 3, 9)

from .plan import FFTPlan   # noqa: E402,F401
from .cuda import Plan      # noqa: E402,F401

__all__ = ["Plan", "FFTPlan", "VERSION"]
