"""Build libb2fft.so in-tree with nvcc for sm_100a:  python -m pyfft_b200.build [-jN] [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")


def build(jobs=None, force=False, verbose=False):
    jobs = jobs or os.cpu_count() or 4
    cmd = ["make", "-C", CSRC, "-j%d" % jobs]
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, stdout=subprocess.DEVNULL)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stdout.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libb2fft.so failed (see output above)")
    lib = os.path.join(HERE, "libb2fft.so")
    if not os.path.exists(lib):
        raise RuntimeError("make succeeded but %s is missing" % lib)
    return lib


if __name__ == "__main__":
    j = None
    for a in sys.argv[1:]:
        if a.startswith("-j") and len(a) > 2:
            j = int(a[2:])
    print(build(j, force="--force" in sys.argv, verbose=True))
