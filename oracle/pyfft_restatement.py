"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference pyfft algorithm.

This is NOT the product and is never called from the product path.  It restates,
in working precision (float32 or float64), what the reference's rendered kernels
compute for one ``plan.execute`` call, so that the CUDA path can be compared with
"what pyfft would have produced" even though pyfft itself cannot run in this image
(Python 2 + Mako + PyCUDA/PyOpenCL; see oracle/__init__.py).

What follows what (all paths relative to /root/reference):

* planner           pyfft/plan.py:111-171   X local-or-global, Y and Z always global
* local radix table pyfft/kernel_helpers.py:10-65   getRadixArray(n, 0)
* global radix info pyfft/kernel_helpers.py:67-122  getGlobalRadixInfo(n)
* butterflies       pyfft/kernel.mako:93-214        fftKernel2/4/8/16 (+ their output
                                                    permutations, so outputs are natural order)
* local twiddle     pyfft/kernel.mako:566-597       ang = (scalar)(2*dir*pi*k/data_len) * (scalar)m
* global twiddles   pyfft/kernel.mako:918-930       ang = (scalar)(2*dir*pi*k/radix) * j
                    pyfft/kernel.mako:957-971       ang1 = (scalar)(2*dir*pi/curr_n) * l; ang = ang1 * (k+idx)
* complex multiply  pyfft/kernel.mako:64            (-a.y*b.y + a.x*b.x, a.y*b.x + a.x*b.y)
* scaling           pyfft/kernel.py:23-37 + kernel.mako:271-278  division by scale_coeff in the last kernel

Data movement (thread/shared-memory index arithmetic, kernel.mako:280-558,599-688)
is not restated: it only permutes values.  Each pass is written as the DIF step it
implements on a state array indexed (K, n_rem):  n_rem = m + M*j  ->  (K + P*k, m).
Rounding differs from a real run only through the device libm (sincos) and FMA
contraction, neither of which the reference pins.
"""
import math

import numpy as np

DIR_FWD = -1   # kernel.mako:3-4  postfix = {1: "Inv", -1: "Fwd"}
DIR_INV = 1


# ----------------------------------------------------------------------------- planner helpers
def log2i(n):
    """kernel_helpers.py:2-8"""
    return int(n).bit_length() - 1


def get_radix_array(n, max_radix=0):
    """kernel_helpers.py:10-65"""
    if max_radix > 1:
        max_radix = min(n, max_radix)
        out = []
        while n > max_radix:
            out.append(max_radix)
            n //= max_radix
        out.append(n)
        return out
    table = {2: [2], 4: [4], 8: [8], 16: [8, 2], 32: [8, 4], 64: [8, 8], 128: [8, 4, 4],
             256: [4, 4, 4, 4], 512: [8, 8, 8], 1024: [16, 16, 4], 2048: [8, 8, 8, 4]}
    if n not in table:
        raise Exception("Wrong problem size: " + str(n))
    return list(table[n])


def get_global_radix_info(n):
    """kernel_helpers.py:67-122"""
    base_radix = min(n, 128)
    radix = []
    N = n
    while N > base_radix:
        N //= base_radix
        radix.append(base_radix)
    radix.append(N)
    R1, R2 = [], []
    for B in radix:
        if B <= 8:
            R1.append(B)
            R2.append(1)
        else:
            r1 = 2
            r2 = B // r1
            while r2 > r1:
                r1 *= 2
                r2 = B // r1
            R1.append(r1)
            R2.append(r2)
    return radix, R1, R2


def max_smem_fft_size(scalar):
    """plan.py:32,46"""
    return 2048 if scalar == np.float32 else 1024


def kernel_chain(x, y, z, scalar):
    """List of ('local'|'global', axis, n) the reference would launch (plan.py:111-171)."""
    chain = []
    if x > max_smem_fft_size(scalar):
        for p in range(len(get_global_radix_info(x)[0])):
            chain.append(("global", 0, x, p))
    elif x > 1:
        chain.append(("local", 0, x, 0))
    if y > 1:
        for p in range(len(get_global_radix_info(y)[0])):
            chain.append(("global", 1, y, p))
    if z > 1:
        for p in range(len(get_global_radix_info(z)[0])):
            chain.append(("global", 2, z, p))
    return chain


# ----------------------------------------------------------------------------- register arithmetic
class C:
    """A 'complex register' holding one value per butterfly: (x, y) real arrays."""
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x = x
        self.y = y


def _add(a, b):
    return C(a.x + b.x, a.y + b.y)


def _sub(a, b):
    return C(a.x - b.x, a.y - b.y)


def _cmul(a, b):
    """kernel.mako:64 complex_mul"""
    return C(-a.y * b.y + a.x * b.x, a.y * b.x + a.x * b.y)


def _ctm(a, d):
    """kernel.mako:68 conj_transp_and_mul(a, b) = (-a.y*b, a.x*b)"""
    return C(-a.y * d, a.x * d)


def _k2s(a, i, j):
    """kernel.mako:102-109 fftKernel2S"""
    c = a[i]
    a[i] = _add(c, a[j])
    a[j] = _sub(c, a[j])


def _const(scalar, x, y):
    return C(scalar(x), scalar(y))


def fft_kernel2(a, d, scalar):
    """kernel.mako:93-100"""
    _k2s(a, 0, 1)


def fft_kernel4(a, d, scalar, o=(0, 1, 2, 3)):
    """kernel.mako:111-134 (fftKernel4 / fftKernel4s on registers o[0..3])"""
    _k2s(a, o[0], o[2])
    _k2s(a, o[1], o[3])
    _k2s(a, o[0], o[1])
    a[o[3]] = _ctm(a[o[3]], scalar(d))
    _k2s(a, o[2], o[3])
    a[o[1]], a[o[2]] = a[o[2]], a[o[1]]


def fft_kernel8(a, d, scalar):
    """kernel.mako:136-166"""
    s = math.sin(math.pi / 4)
    w1 = _const(scalar, s, s * d)
    w3 = _const(scalar, -s, s * d)
    for i in range(4):
        _k2s(a, i, i + 4)
    a[5] = _cmul(w1, a[5])
    a[6] = _ctm(a[6], scalar(d))
    a[7] = _cmul(w3, a[7])
    _k2s(a, 0, 2)
    _k2s(a, 1, 3)
    _k2s(a, 4, 6)
    _k2s(a, 5, 7)
    a[3] = _ctm(a[3], scalar(d))
    a[7] = _ctm(a[7], scalar(d))
    _k2s(a, 0, 1)
    _k2s(a, 2, 3)
    _k2s(a, 4, 5)
    _k2s(a, 6, 7)
    a[1], a[4] = a[4], a[1]          # bitreverse8
    a[3], a[6] = a[6], a[3]


def fft_kernel16(a, d, scalar):
    """kernel.mako:168-214"""
    w0 = scalar(math.cos(math.pi / 8))
    w1 = scalar(math.sin(math.pi / 8))
    w2 = scalar(math.sin(math.pi / 4))
    sd = scalar(d)
    for i in range(4):
        fft_kernel4(a, d, scalar, (i, i + 4, i + 8, i + 12))
    a[5] = _cmul(a[5], C(w0, sd * w1))
    a[7] = _cmul(a[7], C(w1, sd * w0))
    t = C(w2, sd * w2)
    a[6] = _cmul(a[6], t)
    a[9] = _cmul(a[9], t)
    a[10] = _ctm(a[10], sd)
    t = C(-w2, sd * w2)
    a[11] = _cmul(a[11], t)
    a[14] = _cmul(a[14], t)
    a[13] = _cmul(a[13], C(w1, sd * w0))
    a[15] = _cmul(a[15], C(-w0, -sd * w1))
    for b in range(0, 16, 4):
        fft_kernel4(a, d, scalar, (b, b + 1, b + 2, b + 3))
    for i, j in ((1, 4), (2, 8), (3, 12), (6, 9), (7, 13), (11, 14)):   # bitreverse4x4
        a[i], a[j] = a[j], a[i]


_KERNELS = {2: fft_kernel2, 4: fft_kernel4, 8: fft_kernel8, 16: fft_kernel16}


def _butterfly(regs, radix, d, scalar):
    """Radix-`radix` register FFT, natural-order output (radix 1 = identity)."""
    if radix == 1:
        return
    _KERNELS[radix](regs, d, scalar)


def _exp(ang):
    """kernel.mako:35-44 complex_exp: w = (cos ang, sin ang) in working precision."""
    return C(np.cos(ang), np.sin(ang))


# ----------------------------------------------------------------------------- passes
# State convention for one axis: arrays re/im of shape (B, P, L): B independent lines,
# P = product of the radices already done (index K), L = remaining length (index n_rem).

def _split_regs(re, im, R):
    B, P, L = re.shape
    M = L // R
    r4 = re.reshape(B, P, R, M)
    i4 = im.reshape(B, P, R, M)
    return [C(r4[:, :, j, :], i4[:, :, j, :]) for j in range(R)], M


def _merge_regs(regs, B, P, R, M):
    # new K' = K + P*k  ->  axis order (k, K)
    re = np.stack([r.x for r in regs], axis=1).reshape(B, R * P, M)
    im = np.stack([r.y for r in regs], axis=1).reshape(B, R * P, M)
    return re, im


def local_fft_lines(re, im, d, scalar):
    """One localKernel launch (kernel.mako:725-803) on lines of length n = re.shape[-1]."""
    B, n = re.shape
    radix_arr = get_radix_array(n, 0)
    re = re.reshape(B, 1, n)
    im = im.reshape(B, 1, n)
    data_len = n
    for r, R in enumerate(radix_arr):
        P = re.shape[1]
        regs, M = _split_regs(re, im, R)
        _butterfly(regs, R, d, scalar)
        if r < len(radix_arr) - 1:
            angf = np.arange(M).astype(scalar)                      # kernel.mako:574-586
            for k in range(1, R):
                ang = scalar(2 * d * math.pi * k / data_len) * angf  # kernel.mako:591
                regs[k] = _cmul(regs[k], _exp(ang))
            data_len //= R
        re, im = _merge_regs(regs, B, P, R, M)
    return re.reshape(B, n), im.reshape(B, n)


def global_pass_lines(re, im, n, pass_num, d, scalar):
    """One globalKernel launch (kernel.mako:805-1047) = pass `pass_num` of the chain for length n.

    re/im: (B, P, L) state with L = curr_n for this pass.
    """
    radix_arr, r1_arr, r2_arr = get_global_radix_info(n)
    radix, R1, R2 = radix_arr[pass_num], r1_arr[pass_num], r2_arr[pass_num]
    B, P, L = re.shape
    curr_n = L
    M = L // radix
    # n_rem = m + M*jj, jj = a*R2 + j_thr   (kernel.mako:892-912)
    r5 = re.reshape(B, P, R1, R2, M)
    i5 = im.reshape(B, P, R1, R2, M)
    regs = [C(r5[:, :, a], i5[:, :, a]) for a in range(R1)]          # each (B, P, R2, M)
    _butterfly(regs, R1, d, scalar)                                   # kernel.mako:914
    if R2 > 1:
        j = np.arange(R2).astype(scalar).reshape(1, 1, R2, 1)
        for k in range(1, R1):
            ang = scalar(2 * d * math.pi * k / radix) * j             # kernel.mako:926
            regs[k] = _cmul(regs[k], _exp(ang))
        # shuffle (kernel.mako:932-949): R2-point FFT over j_thr for every k1
        out = [None] * radix                                          # index kk = k1 + R1*k2
        for k1 in range(R1):
            sub = [C(regs[k1].x[:, :, t, :], regs[k1].y[:, :, t, :]) for t in range(R2)]
            _butterfly(sub, R2, d, scalar)                            # kernel.mako:951-953
            for k2 in range(R2):
                out[k1 + R1 * k2] = sub[k2]
    else:
        out = [C(r.x[:, :, 0, :], r.y[:, :, 0, :]) for r in regs]
    if pass_num < len(radix_arr) - 1:                                 # kernel.mako:957-971
        l = np.arange(M).astype(scalar).reshape(1, 1, M)
        ang1 = scalar(2 * d * math.pi / curr_n) * l
        for kk in range(radix):
            ang = ang1 * scalar(kk)
            out[kk] = _cmul(out[kk], _exp(ang))
    return _merge_regs(out, B, P, radix, M)


def _axis_to_lines(arr, axis):
    moved = np.moveaxis(arr, axis, -1)
    return np.ascontiguousarray(moved).reshape(-1, moved.shape[-1]), moved.shape


def _lines_to_axis(lines, moved_shape, axis):
    return np.moveaxis(lines.reshape(moved_shape), -1, axis)


def pyfft_execute(data_re, data_im, shape, batch=1, inverse=False, normalize=True, scale=1.0):
    """Restated ``FFTPlan.execute`` (plan.py:173-284).

    data_re/data_im: real arrays (float32 or float64) holding ``batch`` transforms of
    numpy-order ``shape`` back to back.  Returns (re, im) of the same dtype/shape.
    """
    scalar = data_re.dtype.type
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    shape = tuple(int(s) for s in shape)
    if not 1 <= len(shape) <= 3:
        raise ValueError("Wrong shape")
    xyz = tuple(reversed(shape)) + (1,) * (3 - len(shape))            # plan.py:73-89
    x, y, z = xyz
    size = x * y * z
    if 2 ** log2i(size) != size:
        raise ValueError("Array dimensions must be powers of two")    # plan.py:23-24
    d = DIR_INV if inverse else DIR_FWD
    re = np.array(data_re, dtype=scalar).reshape(batch, z, y, x)
    im = np.array(data_im, dtype=scalar).reshape(batch, z, y, x)

    # X (plan.py:140-155), Y (160-163), Z (164-167)
    for axis_id, n in ((0, x), (1, y), (2, z)):
        if n <= 1:
            continue
        np_axis = 3 - axis_id
        lr, mshape = _axis_to_lines(re, np_axis)
        li, _ = _axis_to_lines(im, np_axis)
        if axis_id == 0 and n <= max_smem_fft_size(scalar):
            lr, li = local_fft_lines(lr, li, d, scalar)
        else:
            B = lr.shape[0]
            sr, si = lr.reshape(B, 1, n), li.reshape(B, 1, n)
            for p in range(len(get_global_radix_info(n)[0])):
                sr, si = global_pass_lines(sr, si, n, p, d, scalar)
            lr, li = sr.reshape(B, n), si.reshape(B, n)
        re = _lines_to_axis(lr, mshape, np_axis)
        im = _lines_to_axis(li, mshape, np_axis)

    # scaling, fused into the last kernel's stores as a division (kernel.py:23-37)
    if d == DIR_FWD:
        coeff = 1 if scale == 1.0 else 1.0 / scale
    else:
        coeff = (size if normalize else 1.0) * scale
    if coeff != 1:
        re = re / scalar(coeff)
        im = im / scalar(coeff)
    out_shape = np.asarray(data_re).shape
    return (np.ascontiguousarray(re).reshape(out_shape).astype(scalar),
            np.ascontiguousarray(im).reshape(out_shape).astype(scalar))


def pyfft_execute_complex(data, shape, batch=1, inverse=False, normalize=True, scale=1.0):
    """Interleaved-layout convenience wrapper (plan.py:261-271)."""
    data = np.asarray(data)
    scalar = np.float32 if data.dtype == np.complex64 else np.float64
    re, im = pyfft_execute(np.ascontiguousarray(data.real).astype(scalar),
                           np.ascontiguousarray(data.imag).astype(scalar),
                           shape, batch, inverse, normalize, scale)
    return (re + 1j * im).astype(data.dtype)
