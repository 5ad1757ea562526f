"""SASS evidence for every kernel of libskfem_b200.so: per-kernel mnemonic histogram
(cuobjdump -sass, sm_100a) with the instructions that prove the B200 features used:
UBLKCP (TMA bulk copy), SYNCS (mbarrier), LDGSTS (cp.async), DMMA (FP64 tensor core),
DFMA/DADD/DMUL (FP64 pipe), BAR, SHFL.

    python tools/sass_summary.py r1      -> profiles/r1_sass_all_kernels.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "scikit-fem_b200", "skfem_b200", "libskfem_b200.so")
KEY = ["UBLKCP", "SYNCS", "LDGSTS", "DMMA", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG",
       "STG", "BAR", "SHFL", "ATOM", "RED"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True,
                               text=True).stdout.splitlines()
    out = ["# SASS mnemonic counts per kernel of libskfem_b200.so (cuobjdump -sass, sm_100a, "
           "nvcc -O3 -fmad=false)", "# columns: total instructions | " + " ".join(KEY), ""]
    for (mangled, hist), dem in zip(kernels.items(), demangled):
        total = sum(hist.values())
        if total == 0:
            continue
        out.append(dem)
        out.append("    {:6d} | ".format(total) + " ".join(
            "{}={}".format(k, hist[k]) for k in KEY if hist[k]))
        top = ", ".join("{} {}".format(k, v) for k, v in hist.most_common(8))
        out.append("           top: " + top)
    path = os.path.join(ROOT, "profiles", tag + "_sass_all_kernels.txt")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", path, "kernels:", len([k for k in kernels.values() if sum(k.values())]))


if __name__ == "__main__":
    main()
