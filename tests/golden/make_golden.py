"""Generates tests/golden/*.npz: seeded inputs + expected outputs of the hot path.

The reference itself cannot be imported here (Python 2 + Mako + PyCUDA/PyOpenCL, see
oracle/__init__.py), so the expected values come from the two oracles:
  * ``expect64``  float64 numpy.fft (what the reference's own tests compare against,
                  test/test_errors.py:35-38,105-112)
  * ``pyfft32/64`` the numpy restatement of the reference's algorithm in working precision
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import numpy_oracle as no            # noqa: E402
from oracle import pyfft_restatement as pr       # noqa: E402

CASES = [
    # name, shape, batch, dtype, inverse, normalize, scale
    ("cfg1_fwd", (1024,), 16, np.complex64, False, True, 1.0),
    ("cfg1_inv", (1024,), 16, np.complex64, True, True, 1.0),
    ("cfg1_inv_nonorm", (1024,), 4, np.complex64, True, False, 1.0),
    ("n4096_b2_fwd", (4096,), 2, np.complex64, False, True, 1.0),
    ("n4096_b2_split_inv", (4096,), 2, np.float32, True, True, 1.0),
    ("d2_64x128_fwd", (64, 128), 3, np.complex64, False, True, 1.0),
    ("d3_16x32x64_dp_fwd", (16, 32, 64), 1, np.complex128, False, True, 1.0),
    ("d3_16x16x16_dp_split_inv", (16, 16, 16), 2, np.float64, True, True, 1.0),
    ("n16_scale10_fwd", (16,), 1, np.complex64, False, True, 10.0),
    ("n8192_fwd", (8192,), 1, np.complex64, False, True, 1.0),
    # axes longer than one CTA holds: the four-step path here, the reference's global-kernel chains there
    ("n32768_fwd", (32768,), 1, np.complex64, False, True, 1.0),
    ("d2_4096x8_inv", (4096, 8), 1, np.complex64, True, True, 1.0),
    # round 2 kernels: short rows (16-byte loads + warp shuffles), long rows (staging slot = exchange buffer), a 2048-long
    # strided axis (streamed fused two-step kernel)
    ("n16_b64_split_fwd", (16,), 64, np.float32, False, True, 1.0),
    ("n32_b8_inv", (32,), 8, np.complex64, True, True, 1.0),
    ("n16384_fwd", (16384,), 1, np.complex64, False, True, 1.0),
    ("n8192_dp_inv", (8192,), 1, np.complex128, True, True, 1.0),
    ("d2_2048x16_fwd", (2048, 16), 1, np.complex64, False, True, 1.0),
]


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    force = "--force" in sys.argv
    for i, (name, shape, batch, dtype, inverse, normalize, scale) in enumerate(CASES):
        if os.path.exists(os.path.join(out_dir, name + ".npz")) and not force:
            continue                                  # committed fixtures are only rewritten with --force
        data = no.make_input(shape, batch, dtype, seed=2000 + i)
        if isinstance(data, tuple):
            re, im = data
        else:
            re, im = np.ascontiguousarray(data.real), np.ascontiguousarray(data.imag)
        z = re.astype(np.float64) + 1j * im.astype(np.float64)
        expect64 = no.fft_oracle(z, shape, batch, inverse, normalize, scale)
        pre, pim = pr.pyfft_execute(re, im, shape, batch, inverse, normalize, scale)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), re=re, im=im, expect64=expect64,
                            pyfft_re=pre, pyfft_im=pim, shape=np.array(shape), batch=batch,
                            dtype=np.dtype(dtype).str, inverse=inverse, normalize=normalize, scale=scale)
        print(name, "rel-L2 restatement vs numpy: %.3g" % no.rel_l2(pre + 1j * pim, expect64))


if __name__ == "__main__":
    main()
