#!/usr/bin/env python
"""Per-kernel SASS evidence for libb2fft.so (no GPU needed): `cuobjdump -sass` split by function, demangled, with
the instruction counts that show what the kernels are made of -- packed FP32 (FFMA2/FADD2/FMUL2), FP64 (DFMA/DADD/
DMUL), TMA (UTMALDG / UTMASTG tensor copies, UBLKCP bulk copies), mbarrier traffic (SYNCS), global / shared memory
accesses, L2 policy instructions of the fused two-step kernels, and the absence of warp shuffles.

    python tools/sass_summary.py [--lib pyfft_b200/libb2fft.so] [--out profiles/sass_summary.txt]
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "DADD", "DMUL", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS",
        "SHFL", "BAR", "CCTL", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "UCGABAR"]


def demangle(names):
    for tool in ("cu++filt", "c++filt"):
        try:
            out = subprocess.run([tool], input="\n".join(names), stdout=subprocess.PIPE, text=True, check=True).stdout
            return out.splitlines()
        except Exception:
            continue
    return names


def shorten(name):
    name = re.sub(r"\bb2::", "", name)
    name = re.sub(r"\(PassParams<.*$", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?:int|bool)\)", "", name)
    name = re.sub(r"TileCfg<(\w+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+)>", r"Cfg<\1,n\2,w\3,g\4,r\5x\6x\7x\8,tw\9>", name)
    return name[:150]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "pyfft_b200", "libb2fft.so"))
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "sass_summary.txt"))
    args = ap.parse_args()
    proc = subprocess.Popen(["cuobjdump", "-sass", args.lib], stdout=subprocess.PIPE, text=True, bufsize=1 << 20)
    funcs, cur, arch = collections.OrderedDict(), None, set()
    op_re = re.compile(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    for line in proc.stdout:
        if line.startswith("\t\tFunction : ") or "Function : " in line[:20]:
            cur = line.split("Function : ")[1].strip()
            funcs[cur] = collections.Counter()
            continue
        if "arch = sm_" in line:
            arch.add(line.split("arch = ")[1].strip())
            continue
        if cur is None:
            continue
        m = op_re.match(line)
        if not m:
            continue
        op = m.group(1)
        c = funcs[cur]
        c["_total"] += 1
        base = op.split(".")[0]
        c[base] += 1
        if op.startswith("ST.E") and ".SYS" in op:
            c["ST.E.STRONG.SYS"] += 1
        if op.startswith("LD.E") and ".SYS" in op:
            c["LD.E.STRONG.SYS"] += 1
    proc.wait()
    names = list(funcs)
    pretty = dict(zip(names, demangle(names)))
    fam = collections.OrderedDict()
    lines = []
    for n in names:
        c = funcs[n]
        short = shorten(pretty[n])
        family = short.split("<")[0]
        f = fam.setdefault(family, collections.Counter())
        f["_kernels"] += 1
        for k in COLS + ["_total"]:
            f[k] += c[k]
        lines.append("%-150s %7d " % (short, c["_total"]) + " ".join("%6d" % c[k] for k in COLS))
    with open(args.out, "w") as out:
        out.write("SASS summary of %s (cuobjdump -sass; arch %s; %d kernels)\n" % (os.path.relpath(args.lib, ROOT), ",".join(sorted(arch)) or "?", len(names)))
        out.write("Blackwell-native evidence: packed FP32 math (FFMA2/FADD2/FMUL2), TMA tensor copies (UTMALDG/UTMASTG), TMA bulk copies (UBLKCP),\n"
                  "mbarrier traffic (SYNCS), asynchronous global->shared copies (LDGSTS: the in-place refill of fused2p_fft_kernel); exchanges go through\n"
                  "shared memory (SHFL only in the lane-pair tuning variant fused2w_fft_kernel); the fused two-step kernels carry L2 cache-control\n"
                  "instructions (CCTL = discard.L2 / prefetch.L2); slab_signal/wait are the .SYS-scope flag kernels.\n\n")
        hdr = "%-150s %7s " % ("kernel family (sum over its instantiations)", "instrs") + " ".join("%6s" % k[:6] for k in COLS)
        out.write(hdr + "\n")
        for family, f in fam.items():
            out.write("%-150s %7d " % ("%s  x%d" % (family, f["_kernels"]), f["_total"]) + " ".join("%6d" % f[k] for k in COLS) + "\n")
        out.write("\n" + "%-150s %7s " % ("kernel", "instrs") + " ".join("%6s" % k[:6] for k in COLS) + "\n")
        out.write("\n".join(lines) + "\n")
    print("wrote %s: %d kernels, %d families" % (args.out, len(names), len(fam)))


if __name__ == "__main__":
    sys.exit(main())
