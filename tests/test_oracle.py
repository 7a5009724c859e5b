"""CPU tests of the oracle itself: the restatement of the reference's algorithm is pinned
against every known-answer test the reference's own suite holds for this path and against its
parity criterion (agreement with numpy.fft under the reference's epsilon)."""
import glob
import os

import numpy as np
import pytest

from oracle import numpy_oracle as no
from oracle import pyfft_restatement as pr

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("normalize", [True, False])
def test_kat_ones16_normalize(dtype, normalize):
    """reference test/test_functionality.py:53-77"""
    data = np.ones(16, dtype=dtype)
    fw = pr.pyfft_execute_complex(data, (16,), normalize=normalize)
    assert np.sum(np.abs(np.fft.fft(data) - fw)) / data.size < 1e-6
    back = pr.pyfft_execute_complex(fw, (16,), inverse=True, normalize=normalize)
    coeff = 1 if normalize else data.size
    assert np.sum(np.abs(data * coeff - back)) / data.size < 1e-6


@pytest.mark.parametrize("scale", [1.0, 10.0])
def test_kat_scale(scale):
    """reference test/test_functionality.py:79-100"""
    data = np.ones(16, dtype=np.complex64)
    fw = pr.pyfft_execute_complex(data, (16,), scale=scale)
    assert np.sum(np.abs(np.fft.fft(data) * scale - fw)) / data.size < 1e-6
    back = pr.pyfft_execute_complex(fw, (16,), inverse=True, scale=scale)
    assert np.sum(np.abs(data - back)) / data.size < 1e-6


def test_kat_ones8192_roundtrip():
    """reference test/test_functionality.py:102-115 (the global-kernel 1D path)"""
    data = np.ones(8192, dtype=np.complex64)
    fw = pr.pyfft_execute_complex(data, (8192,))
    back = pr.pyfft_execute_complex(fw, (8192,), inverse=True)
    assert np.sum(np.abs(data - back)) / data.size < 1e-6


def test_kat_doc_example_16x16():
    """reference doc/source/index.rst:61-99, examples/cuda_basic.py:16-28"""
    data = np.ones((16, 16), dtype=np.complex64)
    fw = pr.pyfft_execute_complex(data, (16, 16))
    assert abs(fw[0, 0] - 256) < 1e-4
    assert np.abs(fw).sum() - abs(fw[0, 0]) < 1e-3
    back = pr.pyfft_execute_complex(fw, (16, 16), inverse=True)
    assert np.abs(back - data).sum() / data.size < 1e-6


def test_radix_tables():
    """reference pyfft/kernel_helpers.py:10-122 (values quoted in SURVEY.md section 8a)"""
    assert pr.get_radix_array(1024) == [16, 16, 4]
    assert pr.get_radix_array(2048) == [8, 8, 8, 4]
    assert pr.get_radix_array(256) == [4, 4, 4, 4]
    assert pr.get_radix_array(32) == [8, 4]
    assert pr.get_radix_array(4096, 16) == [16, 16, 16]
    assert pr.get_global_radix_info(4096) == ([128, 32], [16, 8], [8, 4])
    assert pr.get_global_radix_info(1024) == ([128, 8], [16, 8], [8, 1])
    assert pr.get_global_radix_info(256) == ([128, 2], [16, 2], [8, 1])
    assert pr.get_global_radix_info(2048) == ([128, 16], [16, 4], [8, 4])
    assert [k[0] for k in pr.kernel_chain(4096, 1, 1, np.float32)] == ["global", "global"]
    assert len(pr.kernel_chain(1024, 1024, 1, np.float32)) == 3
    assert len(pr.kernel_chain(256, 256, 256, np.float64)) == 5


# the reference's parity grid (test/test_errors.py:122-140), at sizes the numpy restatement
# finishes in seconds
GRID = [((8,), 16), ((256,), 16), ((512,), 4), ((1024,), 16), ((2048,), 4), ((8192,), 2), ((1 << 16,), 1),
        ((16, 16), 16), ((128, 16), 2), ((16, 128), 2), ((256, 256), 1), ((1024, 16), 1),
        ((16, 16, 16), 2), ((16, 128, 16), 1), ((128, 16, 16), 1), ((16, 16, 128), 1)]


@pytest.mark.parametrize("shape,batch", GRID)
@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.float64, np.complex128])
def test_reference_parity_criterion(shape, batch, dtype):
    """The restated reference satisfies its own acceptance test: forward ~ numpy.fft.fftn and
    forward->inverse ~ input under eps 1.1e-6 / 1e-11 (test/test_errors.py:20-23,105-112)."""
    data = no.make_input(shape, batch, dtype, seed=7)
    if isinstance(data, tuple):
        re, im = data
    else:
        re, im = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
    z = re + 1j * im
    fre, fim = pr.pyfft_execute(re, im, shape, batch)
    eps = no.reference_epsilon(dtype)
    ref = no.fft_oracle(z, shape, batch)
    assert no.pyfft_difference(ref, fre + 1j * fim, batch) < eps
    assert no.rel_l2(fre + 1j * fim, ref) < no.tolerance(dtype, int(np.prod(shape)))
    bre, bim = pr.pyfft_execute(fre, fim, shape, batch, inverse=True)
    assert no.pyfft_difference(z, bre + 1j * bim, batch) < eps


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures_match_oracles(path):
    """Committed fixtures are reproducible from both oracles (guards against silent drift)."""
    g = np.load(path)
    shape, batch = tuple(int(s) for s in g["shape"]), int(g["batch"])
    inverse, normalize, scale = bool(g["inverse"]), bool(g["normalize"]), float(g["scale"])
    z = g["re"].astype(np.float64) + 1j * g["im"].astype(np.float64)
    want = no.fft_oracle(z, shape, batch, inverse, normalize, scale)
    assert no.rel_l2(g["expect64"], want) < 1e-14
    pre, pim = pr.pyfft_execute(g["re"], g["im"], shape, batch, inverse, normalize, scale)
    tol = no.tolerance(g["re"].dtype, int(np.prod(shape)))
    assert no.rel_l2(pre + 1j * pim, g["pyfft_re"] + 1j * g["pyfft_im"]) < 1e-6 * tol + 1e-12
    assert no.rel_l2(g["pyfft_re"] + 1j * g["pyfft_im"], want) < tol


def test_wrong_sizes_raise():
    """reference pyfft/plan.py:23-24,87-89"""
    with pytest.raises(ValueError):
        pr.pyfft_execute(np.zeros(17, np.float32), np.zeros(17, np.float32), (17,))
    with pytest.raises(ValueError):
        pr.pyfft_execute(np.zeros(16, np.float32), np.zeros(16, np.float32), (2, 2, 2, 2))
