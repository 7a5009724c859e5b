"""CPU tests of the host planner (no GPU): b2fft_plan_preview builds the pass list exactly as
b2fft_plan_create does.  One pass per axis while the axis fits one CTA; longer axes are split
four-step style (the counterpart of the reference's local-vs-global decision, pyfft/plan.py:141-167,
and its base-128 chains, pyfft/kernel_helpers.py:67-122)."""
import ctypes
import re

import pytest


def _preview(lib, x, y=1, z=1, prec=0, layout=0, mask=7):
    buf = ctypes.create_string_buffer(4096)
    from pyfft_b200 import _lib
    _lib.check(lib.b2fft_plan_preview((ctypes.c_int64 * 3)(x, y, z), mask, prec, layout, buf, len(buf)))
    out = []
    for line in buf.value.decode().splitlines():
        m = re.match(r"axis=(\w) n=(\d+) inner=(\d+) variant=(\S+)(?: fs=(\d+)x(\d+))?", line)
        assert m, line
        out.append({"axis": m.group(1), "n": int(m.group(2)), "inner": int(m.group(3)), "variant": m.group(4),
                    "fs": (int(m.group(5)), int(m.group(6))) if m.group(5) else None})
    return out


def test_one_pass_per_axis_when_it_fits(built_lib):
    for prec, xmax in ((0, 1 << 14), (1, 1 << 13)):
        p = _preview(built_lib, xmax, 2048, 2048, prec=prec)
        assert [(q["axis"], q["n"], q["inner"]) for q in p] == [("X", xmax, 1), ("Y", 2048, xmax), ("Z", 2048, xmax * 2048)]
        assert all(q["fs"] is None for q in p)
    assert _preview(built_lib, 1, 1, 1) == []                                    # nothing to do
    assert [q["axis"] for q in _preview(built_lib, 8, 1, 16)] == ["X", "Z"]       # size-1 axes are no-ops
    assert [q["axis"] for q in _preview(built_lib, 64, 64, 64, mask=2)] == ["Y"]  # axis masks (slab plans)


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("lg", [14, 15, 16, 20, 22, 23, 24, 27])
def test_long_contiguous_axis_is_split(built_lib, prec, lg):
    n = 1 << lg
    p = _preview(built_lib, n, prec=prec)
    if lg <= (14 if prec == 0 else 13):
        assert len(p) == 1 and p[0]["fs"] is None
        return
    assert len(p) == (2 if lg <= 22 else 3)
    # the factors multiply to n; every transposing pass carries the twiddle of the axis it splits
    prod, inner, remaining = 1, 1, n
    for q in p[:-1]:
        n1, n2 = q["fs"]
        assert q["n"] == n1 and n1 * n2 == remaining and q["inner"] == n2 * inner
        prod, inner, remaining = prod * n1, inner * n1, n2
    assert p[-1]["fs"] is None and p[-1]["n"] == remaining and p[-1]["inner"] == inner
    assert prod * p[-1]["n"] == n
    assert all(q["n"] <= 2048 for q in p)


def test_long_strided_axis_is_split(built_lib):
    p = _preview(built_lib, 64, 4096, 1)
    assert [(q["axis"], q["n"], q["inner"], q["fs"]) for q in p] == [("X", 64, 1, None), ("Y", 64, 64 * 64, (64, 64)),
                                                                  ("Y", 64, 64 * 64, None)]
    p = _preview(built_lib, 4096, 8, 4096)
    assert [q["axis"] for q in p] == ["X", "Y", "Z", "Z"] and p[2]["fs"] == (64, 64)


def test_layout_and_pitch_aware_variant_choice(built_lib):
    # split-layout rows take the TMA-staged kernel, interleaved rows the plain high-occupancy one
    assert "_tma" in _preview(built_lib, 4096, layout=1)[0]["variant"]
    assert "_tma" not in _preview(built_lib, 4096, layout=0)[0]["variant"]
    # strided axes: TMA tensor staging for N >= 1024 while rows are < 256 KiB apart ...
    assert "_tmac" in _preview(built_lib, 1024, 1024)[1]["variant"]
    assert "_tmac" in _preview(built_lib, 2048, 2048, 2048, layout=1)[1]["variant"]
    # ... and the fused two-step kernel (128-byte pieces on both DRAM sides) beyond, in both precisions; complex64 axes of
    # length 2048 take its streamed form (fused2p) at every pitch, complex64 1024 beyond 256 KiB
    y, z = _preview(built_lib, 2048, 2048, 2048)[1:3]
    assert y["variant"].endswith("_fused2p") and z["variant"].endswith("_fused2p") and "_w16_" in z["variant"]
    assert _preview(built_lib, 1024, 1024, 1024)[2]["variant"].endswith("_fused2p")
    assert "_fused2" not in _preview(built_lib, 8, 2048, 2048)[1]["variant"]         # inner stride 8: no 16-column tile
    assert "_fused2" in _preview(built_lib, 256, 64, 1024, prec=1)[2]["variant"]
    # ... unless the layout is split or the inner stride is not a multiple of 128 bytes: widest plain tile
    z = _preview(built_lib, 2048, 2048, 2048, layout=1)[2]
    assert "_fused2" not in z["variant"] and "_tmac" not in z["variant"] and "_w8_" in z["variant"]
    assert "_fused2" not in _preview(built_lib, 8, 8192, 2048)[2]["variant"]


def test_short_rows_take_the_shuffle_kernel(built_lib):
    # complex64 / split float32 rows of 4 .. 32 elements: 16-byte accesses + warp-shuffle exchanges (row_shfl_kernel) ...
    for n in (4, 8, 16):
        assert _preview(built_lib, n)[0]["variant"] == "float_n%d_w1_shfl" % (n.bit_length() - 1)
        assert _preview(built_lib, n, layout=1)[0]["variant"].endswith("_shfl")
    for n in (32, 64):           # four elements per lane: 16-byte stores as well, one shuffle stage fewer
        assert _preview(built_lib, n)[0]["variant"] == "float_n%d_w1_shfl4" % (n.bit_length() - 1)
    assert _preview(built_lib, 16, 16, 16)[0]["variant"].endswith("_shfl")        # the X pass of a small 3-D transform
    # ... not for N = 2 and N >= 128 (the tile program is at the copy bandwidth there), not for double precision
    assert "_shfl" not in _preview(built_lib, 2)[0]["variant"] and "_shfl" not in _preview(built_lib, 128)[0]["variant"]
    assert "_shfl" not in _preview(built_lib, 16, prec=1)[0]["variant"]


def test_long_rows_take_the_aliased_tma_kernel(built_lib):
    # rows of 64-128 KiB: persistent TMA-staged kernel whose staging slot is the exchange buffer (tile_fft_kernel_tma_row_alias)
    assert _preview(built_lib, 16384)[0]["variant"].endswith("_tmara")
    assert _preview(built_lib, 8192)[0]["variant"].endswith("_tmara")
    assert _preview(built_lib, 8192, prec=1)[0]["variant"].endswith("_tmara")
    assert _preview(built_lib, 16384, layout=1)[0]["variant"].endswith("_tmara")
    # ... complex128 4096 keeps the plain kernel (two CTAs per SM, 0.95 of the copy bandwidth)
    assert "_tma" not in _preview(built_lib, 4096, prec=1)[0]["variant"]


def test_preview_validation(built_lib):
    from pyfft_b200 import _lib
    buf = ctypes.create_string_buffer(64)
    assert built_lib.b2fft_plan_preview((ctypes.c_int64 * 3)(12, 1, 1), 7, 0, 0, buf, len(buf)) == _lib.E_INVALID
    assert "powers of two" in _lib.last_error()
    assert built_lib.b2fft_plan_preview((ctypes.c_int64 * 3)(16, 1, 1), 7, 3, 0, buf, len(buf)) == _lib.E_INVALID
