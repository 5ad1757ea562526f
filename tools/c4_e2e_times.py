"""Per-call times of the cold public-API call of BASELINE configs[3] (Hex2, 65^3 points):
Mesh + Basis + laplace.assemble -> scipy CSR, with and without the sum-factorised kernel."""
import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import skfem_b200 as fem
from skfem_b200 import form as F
from skfem_b200.models.poisson import laplace
x = np.linspace(0, 1, 65)
m = fem.MeshHex.init_tensor(x, x, x)
p, t = m.p, m.t
def call():
    return laplace.assemble(fem.Basis(fem.MeshHex(p, t), fem.ElementHex2()))
for sf in (True, False, True):
    F.set_options(hex_sumfact=sf)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); A = call(); ts.append(1e3 * (time.perf_counter() - t0))
    print("sumfact", sf, " ".join("%.1f" % v for v in ts), "nnz", A.nnz, flush=True)
pr = cProfile.Profile(); pr.enable(); call(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(18); print(s.getvalue()[:3500])
