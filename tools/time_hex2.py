"""C4 (BASELINE configs[3]: ElementHex2 on MeshHex.init_tensor) timing split: the
element-local kernel (DMMA Gram-matrix contraction, csrc/skb_hex_mma.cu; scalar
kernel with debug flag 8) versus the deterministic CSR reduction.

    python tools/time_hex2.py [--cells 64]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200"))

import torch  # noqa: E402
import skfem_b200 as fem  # noqa: E402
from skfem_b200 import _lib  # noqa: E402
from skfem_b200.models.poisson import laplace, mass  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=64)
ap.add_argument("--scalar", action="store_true", help="also time the scalar kernel (slow)")
args = ap.parse_args()
x = np.linspace(0, 1, args.cells + 1)
b = fem.Basis(fem.MeshHex.init_tensor(x, x, x), fem.ElementHex2())
print("nel", b.nelems, "N", b.N, "nqp", b.nqp)


def timed(fn, reps=5):
    """median of `reps` individually synchronised calls (the 1.5 GB local-data
    allocation of the first calls goes through cudaMalloc)"""
    for _ in range(2):
        r = fn()
        del r
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
        if _ != reps - 1:
            del r
    return float(np.median(ts)), r


for name, form in (("laplace", laplace), ("mass", mass)):
    for flag in ((0, 8) if args.scalar else (0,)):
        _lib.lib().skb_debug_flags(flag)
        ms, loc = timed(lambda: form._local(b))
        flops = b.nelems * 27 * 27 * b.nqp * (2 * 3 + 2 if name == "laplace" else 3)
        print("{:8s} local kernel ({}): {:8.2f} ms  ({:.2f} TFLOP/s of the direct contraction)".format(
            name, "scalar" if flag else "DMMA", ms, flops / ms / 1e9))
        del loc
    _lib.lib().skb_debug_flags(0)
    form.assemble_device(b)
    ms, A = timed(lambda: form.assemble_device(b))
    print("{:8s} warm assemble_device: {:8.2f} ms, nnz {}, {:.3e} el/s".format(
        name, ms, A.nnz, b.nelems / ms * 1e3))
    del A
