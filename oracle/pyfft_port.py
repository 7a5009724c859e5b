"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/pyfft_port.c (the C restatement of the
reference algorithm, OpenMP over lines).  Used by tests as a second, independent checker and by
bench.py as the timed CPU reference arm.  Never imported by the product."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libpyfft_port.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(_PATH)
        for name, ptr in (("pyfft_port_execute_f32", ctypes.c_float), ("pyfft_port_execute_f64", ctypes.c_double)):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_long] * 4 + [ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                                        ctypes.c_int, ctypes.c_int]
        lib.pyfft_port_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def max_threads():
    return int(load().pyfft_port_max_threads())


def execute(data, shape, batch=1, inverse=False, normalize=True, scale=1.0, nthreads=0, out=None):
    """data: complex array (interleaved) or (re, im) tuple of real arrays.  Returns the same kind."""
    lib = load()
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    xyz = tuple(reversed([int(s) for s in shape])) + (1,) * (3 - len(shape))
    if isinstance(data, tuple):
        re = np.ascontiguousarray(data[0])
        im = np.ascontiguousarray(data[1])
        fn = lib.pyfft_port_execute_f32 if re.dtype == np.float32 else lib.pyfft_port_execute_f64
        ore, oim = np.empty_like(re), np.empty_like(im)
        rc = fn(re.ctypes.data, im.ctypes.data, ore.ctypes.data, oim.ctypes.data, xyz[0], xyz[1], xyz[2], batch,
                int(inverse), int(normalize), float(scale), 0, int(nthreads))
        if rc:
            raise ValueError("pyfft_port_execute failed: %d" % rc)
        return ore, oim
    a = np.ascontiguousarray(data)
    fn = lib.pyfft_port_execute_f32 if a.dtype == np.complex64 else lib.pyfft_port_execute_f64
    o = np.empty_like(a) if out is None else out
    rc = fn(a.ctypes.data, None, o.ctypes.data, None, xyz[0], xyz[1], xyz[2], batch, int(inverse), int(normalize),
            float(scale), 1, int(nthreads))
    if rc:
        raise ValueError("pyfft_port_execute failed: %d" % rc)
    return o
