"""CPU test: the exact CUDA thread program (pyfft_b200/csrc/fft_core.cuh) executed on the host,
phase by phase, for EVERY compiled kernel variant (tests/host_emu/emu.cpp), against a long-double
FFT.  Validates index arithmetic, padding, twiddle tables and butterflies without a GPU.  The cases are
built as four translation units in parallel (-DEMU_PART=0..3: rows, strided axes, tuning variants, fused kernels)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_thread_program_on_host(tmp_path):
    src = os.path.join(ROOT, "tests", "host_emu", "emu.cpp")
    procs = []
    for part in range(4):
        exe = str(tmp_path / ("emu%d" % part))
        procs.append((exe, subprocess.Popen(["g++", "-std=c++17", "-O1", "-DEMU_PART=%d" % part, "-I", os.path.join(ROOT, "pyfft_b200", "csrc"),
                                             "-I", os.path.join(ROOT, "tests", "host_emu"), src, "-o", exe])))
    out = ""
    for exe, proc in procs:
        assert proc.wait() == 0, "compilation of %s failed" % exe
    runs = [subprocess.Popen([exe], stdout=subprocess.PIPE, text=True) for exe, _ in procs]
    for run in runs:
        text = run.communicate()[0]
        assert run.returncode == 0, text[-4000:]
        assert "ALL OK" in text
        out += text
    assert out.count(" ok") >= 120 and "FAIL" not in out
