"""Run the other BASELINE configs at (or near) full size on one B200 and check
size-independent properties; prints a markdown table (-> profiles/r1_configs.md).
  C3  ElementVector(ElementTetP2) linear_elasticity, init_tensor `--c3` pts/side (70 = 1.97 M tets)
  C4  MeshHex init_tensor `--c4` cells/side, ElementHex2 laplace + mass (64 = 262 144 hexes)
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import skfem_b200 as fem
from skfem_b200.models.poisson import laplace, mass
from skfem_b200.models.elasticity import linear_elasticity, lame_parameters

ap = argparse.ArgumentParser()
ap.add_argument("--c3", type=int, default=70)
ap.add_argument("--c4", type=int, default=64)
args = ap.parse_args()


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    return r, time.perf_counter() - t0


print("| config | nel | N | nnz | host basis s | cold assemble s | warm assemble_device ms | el/s warm | checks |")
print("|---|---|---|---|---|---|---|---|---|")
if args.c3:
    x = np.linspace(0, 1, args.c3)
    m = fem.MeshTet.init_tensor(x, x, x)
    b, tb = timed(lambda: fem.Basis(m, fem.ElementVector(fem.ElementTetP2())))
    form = linear_elasticity(*lame_parameters(1e3, 0.3))
    A, tc = timed(lambda: form.assemble_device(b))
    _, tw = timed(lambda: form.assemble_device(b))
    # rigid body modes are in the kernel: translations and an infinitesimal rotation
    d = b._dev()
    S = A.to_torch()
    ok = []
    for comp in range(3):
        v = torch.zeros(b.N, dtype=torch.float64, device="cuda"); v[comp::3] = 1.0
        ok.append(float((S @ v).abs().max()))
    scale = float(A.data.abs().max())
    sym = float((S.to_dense() - S.to_dense().T).abs().max()) if b.N < 20000 else float("nan")
    print("| C3 vector-P2 elasticity | %d | %d | %d | %.1f | %.2f | %.1f | %.3g | max|A*translation|/max|A| = %.1e |"
          % (b.nelems, b.N, A.nnz, tb, tc, 1e3 * tw, b.nelems / tw, max(ok) / scale))
    del A, S, b, m
    torch.cuda.empty_cache()
if args.c4:
    x = np.linspace(0, 1, args.c4 + 1)
    m = fem.MeshHex.init_tensor(x, x, x)
    b, tb = timed(lambda: fem.Basis(m, fem.ElementHex2()))
    for name, form in (("laplace", laplace), ("mass", mass)):
        A, tc = timed(lambda: form.assemble_device(b))
        _, tw = timed(lambda: form.assemble_device(b))
        S = A.to_torch()
        one = torch.ones(b.N, dtype=torch.float64, device="cuda")
        r = S @ one
        chk = ("max|A*1|/max|A| = %.1e" % (float(r.abs().max()) / float(A.data.abs().max()))
               if name == "laplace" else "sum(M) = %.15g" % float(r.sum()))
        print("| C4 Hex2 %s | %d | %d | %d (closed form %d) | %.1f | %.2f | %.1f | %.3g | %s |"
              % (name, b.nelems, b.N, A.nnz, (8 * args.c4 + 1) ** 3, tb, tc, 1e3 * tw,
                 b.nelems / tw, chk))
        del A, S
        torch.cuda.empty_cache()
