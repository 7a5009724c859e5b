"""TEST INFRASTRUCTURE ONLY -- CPU oracles for the batched C2C FFT hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm may
import it, and only as the checker (or the thing timed as the CPU baseline),
never as a fallback for the CUDA path.

Contents
--------
numpy_oracle.py        float64 ``numpy.fft`` oracle + the two error metrics
pyfft_restatement.py   numpy restatement of the reference's own algorithm
                       (pyfft/kernel_helpers.py, pyfft/kernel.py, pyfft/kernel.mako)
                       in working precision (fp32 / fp64)
pyfft_port.c           the same algorithm in C with OpenMP over the batch, used
                       as the timed "reference on host cores" arm of bench.py

Parity pinning
--------------
The reference (pyfft 0.3.9) is Python-2 + Mako + PyCUDA/PyOpenCL and cannot be
imported or compiled in this image (no Python 2, no mako, no pyopencl/pocl, no
pycuda), so ``oracle/_ref`` does not exist.  The restatement is pinned against
every known-answer test the reference's own test-suite holds for this path
(test/test_functionality.py:53-115, doc/source/index.rst:61-99) and against the
reference's parity criterion (test/test_errors.py:20-23,105-112: agreement with
``numpy.fft.fftn`` under eps = 1.1e-6 sp / 1e-11 dp).  See tests/test_oracle.py.
"""
