"""CPU test: the exact CUDA thread program (pyfft_b200/csrc/fft_core.cuh) executed on the host,
phase by phase, for EVERY compiled kernel variant (tests/host_emu/emu.cpp), against a long-double
FFT.  Validates index arithmetic, padding, twiddle tables and butterflies without a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_thread_program_on_host(tmp_path):
    exe = str(tmp_path / "emu")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "pyfft_b200", "csrc"),
                    "-I", os.path.join(ROOT, "tests", "host_emu"),
                    os.path.join(ROOT, "tests", "host_emu", "emu.cpp"), "-o", exe], check=True)
    res = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert res.returncode == 0, res.stdout[-4000:]
    assert "ALL OK" in res.stdout
    assert res.stdout.count(" ok") >= 40
