"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and argument validation raises the reference's exception types before any GPU work."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2fft.h")).read()
    return sorted(set(re.findall(r"B2FFT_API[^;]*?\b(b2fft_\w+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    from pyfft_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(built_lib, n), "libb2fft.so does not export " + n
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table and header disagree"


def test_version_and_variants(built_lib):
    assert built_lib.b2fft_version() >= 100
    n = built_lib.b2fft_num_variants()
    assert n >= 40
    buf = ctypes.create_string_buffer(256)
    seen = set()
    for i in range(n):
        assert built_lib.b2fft_variant_info(i, buf, len(buf)) == 0
        f = buf.value.decode().split()
        assert len(f) == 11
        seen.add((int(f[1]), int(f[2]), int(f[3]) == 1))
    # every length 2..2^12 has a contiguous-axis kernel in both precisions
    for prec in (0, 1):
        for lg in range(1, 13):
            assert (prec, lg, True) in seen


def test_invalid_arguments_are_value_errors(built_lib):
    """reference pyfft/plan.py:23-24,48,87-89 / test/test_functionality.py:129-139"""
    from pyfft_b200 import _lib
    h = ctypes.c_void_p()
    dims = (ctypes.c_int64 * 3)(17, 1, 1)
    rc = built_lib.b2fft_plan_create(ctypes.byref(h), 1, dims, 0, 0, 1, 1.0, 1, 0)
    assert rc == _lib.E_INVALID and "powers of two" in _lib.last_error()
    dims = (ctypes.c_int64 * 3)(16, 1, 1)
    assert built_lib.b2fft_plan_create(ctypes.byref(h), 4, dims, 0, 0, 1, 1.0, 1, 0) == _lib.E_INVALID
    assert built_lib.b2fft_plan_create(ctypes.byref(h), 1, dims, 7, 0, 1, 1.0, 1, 0) == _lib.E_INVALID
    assert built_lib.b2fft_plan_create(ctypes.byref(h), 1, dims, 0, 0, 1, 0.0, 1, 0) == _lib.E_INVALID
    assert built_lib.b2fft_execute(None, None, None, None, None, 0, 1, None) == _lib.E_INVALID


def test_slab_plan_argument_validation(built_lib):
    """b2fft_slab_*: bad decompositions are refused with E_INVALID before any device work."""
    from pyfft_b200 import _lib
    h = ctypes.c_void_p()

    def create(dims, rank, nranks, prec=0):
        return built_lib.b2fft_slab_plan_create(ctypes.byref(h), (ctypes.c_int64 * 3)(*dims), prec, 1, 1.0, 1, 0, rank, nranks, 8, 1, 3)
    assert create((64, 64, 64), 0, 3) == _lib.E_INVALID and b"power of two" in built_lib.b2fft_slab_last_error()
    assert create((64, 64, 64), 4, 4) == _lib.E_INVALID
    assert create((64, 64, 48), 0, 4) == _lib.E_INVALID and b"powers of two" in built_lib.b2fft_slab_last_error()
    assert create((2, 64, 64), 0, 4) == _lib.E_INVALID and b"divisible" in built_lib.b2fft_slab_last_error()
    assert create((64, 64, 2), 0, 4) == _lib.E_INVALID
    assert create((64, 64, 64), 0, 2, prec=5) == _lib.E_INVALID
    assert built_lib.b2fft_slab_forward(None, None) == _lib.E_INVALID
    assert built_lib.b2fft_slab_plan_destroy(None) == _lib.OK
    # source-blocked loads are only defined for single-pass contiguous-axis plans
    assert built_lib.b2fft_plan_set_input_blocks(None, 2, None) == _lib.E_INVALID


def test_python_plan_validation_without_gpu():
    from pyfft_b200.cuda import Plan
    with pytest.raises(ValueError):
        Plan((17,), dtype=np.complex64)
    with pytest.raises(ValueError):
        Plan((16,), dtype=np.int32)
    with pytest.raises(ValueError):
        Plan((16, 16, 16, 16), dtype=np.complex64)
    with pytest.raises(ValueError):
        Plan("16", dtype=np.complex64)
    with pytest.raises(ValueError):
        Plan((16, 3), dtype=np.complex64)


def test_no_cpu_fallback():
    """Without a GPU a well-formed plan must fail loudly, not silently compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyfft_b200.cuda import Plan
    with pytest.raises(RuntimeError):
        Plan((16,), dtype=np.complex64)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under pyfft_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pyfft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f
