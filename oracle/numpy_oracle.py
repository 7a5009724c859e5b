"""TEST INFRASTRUCTURE ONLY -- float64 numpy.fft oracle and error metrics.

Transform definition (reference pyfft/kernel.py:23-37,
test/test_functionality.py:53-100):

  forward : Y = scale * fftn(X)           over the last 1..3 axes, never normalised
  inverse : X = ifftn(Y) * size / (scale * (size if normalize else 1))

with ``size = x*y*z`` (not multiplied by batch) and ``batch`` transforms stored
back to back along the leading axis (doc/source/index.rst:254-255).
"""
import numpy as np


def _as_tuple(shape):
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def fft_oracle(data, shape, batch=1, inverse=False, normalize=True, scale=1.0):
    """float64/complex128 result for ``batch`` transforms of ``shape`` stored in ``data``."""
    shape = _as_tuple(shape)
    a = np.asarray(data).astype(np.complex128).reshape((batch,) + shape)
    axes = tuple(range(1, 1 + len(shape)))
    if not inverse:
        res = np.fft.fftn(a, axes=axes) * scale
    else:
        size = int(np.prod(shape))
        res = np.fft.ifftn(a, axes=axes) * size          # un-normalised inverse
        res = res / (scale * (size if normalize else 1))
    return res.reshape(np.asarray(data).shape)


def rel_l2(a, b):
    """BASELINE.json's parity metric: ||a-b||_2 / ||b||_2 (b is the oracle)."""
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def pyfft_difference(arr1, arr2, batch):
    """The reference's own error metric, mean over batch of sum|a-b| / sum|a|
    (test/helpers.py:160-176), without its off-by-one slice."""
    d = np.abs(np.asarray(arr1) - np.asarray(arr2)).reshape(batch, -1).sum(axis=1)
    m = np.abs(np.asarray(arr1)).reshape(batch, -1).sum(axis=1)
    m = np.where(m > 0, m, 1.0)
    return float(np.mean(d / m))


def tolerance(dtype, n_total):
    """north_star tolerance: 1e-5*log2(N) single, 1e-13*log2(N) double, N = x*y*z."""
    lg = max(1.0, float(np.log2(n_total)))
    if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)):
        return 1e-5 * lg
    return 1e-13 * lg


def reference_epsilon(dtype):
    """eps of the reference's own parity test (test/test_errors.py:20-23)."""
    if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)):
        return 1.1e-6
    return 1e-11


def make_input(shape, batch, dtype, seed):
    """Seeded standard-normal test data (same distribution as test/helpers.py:178-196).

    complex dtype -> one interleaved array of shape (batch,)+shape
    real dtype    -> (re, im) pair of that shape (split layout)
    """
    shape = _as_tuple(shape)
    rng = np.random.default_rng(seed)
    full = (batch,) + shape
    re = rng.standard_normal(full)
    im = rng.standard_normal(full)
    dt = np.dtype(dtype)
    if dt.kind == "c":
        fl = np.float32 if dt == np.complex64 else np.float64
        return (re.astype(fl) + 1j * im.astype(fl)).astype(dt)
    return re.astype(dt), im.astype(dt)
