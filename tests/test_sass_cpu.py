"""CPU test: the built library really contains the Blackwell-native instructions DESIGN.md claims
(cuobjdump -sass on the sm_100a cubins inside libb2fft.so; B200_PROFILING.md "what proves a
Blackwell-native kernel"): packed FP32 butterflies (FADD2 / FMUL2 / FFMA2), TMA bulk copies (UBLKCP),
TMA tensor loads / stores (UTMALDG / UTMASTG), mbarrier waits (SYNCS), asynchronous global->shared copies (LDGSTS: the
streamed fused two-step kernel), warp shuffles (SHFL: the short-row kernel) -- and no cuFFT dependency."""
import re
import shutil
import subprocess

import pytest


def test_sass_mnemonics_present(built_lib):
    from pyfft_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        proc = subprocess.Popen([cuobjdump, "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True)
    except OSError:
        pytest.skip("cuobjdump not available")
    counts = {k: 0 for k in ("FADD2", "FMUL2", "FFMA2", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "SHFL")}
    archs = set()
    pat = re.compile(r"\b(FADD2|FMUL2|FFMA2|UBLKCP|UTMALDG|UTMASTG|SYNCS|LDGSTS|SHFL)\b")
    for line in proc.stdout:
        if line.startswith("arch ="):
            archs.add(line.split("=")[1].strip())
        m = pat.search(line)
        if m:
            counts[m.group(1)] += 1
    proc.wait()
    assert archs == {"sm_100a"}, archs
    for k, v in counts.items():
        assert v > 0, "no %s in libb2fft.so" % k
    assert counts["FFMA2"] + counts["FADD2"] + counts["FMUL2"] > 10000


def test_no_cufft_dependency(built_lib):
    from pyfft_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "cufft" not in out.lower() and "cublas" not in out.lower(), out
