"""cProfile of Basis construction for the higher-order configs (host side), on the GPU box."""
import os, sys, cProfile, pstats, io, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import skfem_b200 as fem
which = sys.argv[1] if len(sys.argv) > 1 else "p2"
if which == "p2":
    x = np.linspace(0, 1, 61); m0 = fem.MeshTet.init_tensor(x, x, x); mk = lambda: (fem.MeshTet(m0.p, m0.t), fem.ElementTetP2())
else:
    x = np.linspace(0, 1, 65); m0 = fem.MeshHex.init_tensor(x, x, x); mk = lambda: (fem.MeshHex(m0.p, m0.t), fem.ElementHex2())
def f():
    m, e = mk(); b = fem.Basis(m, e); b._dev(); torch.cuda.synchronize(); return b
f(); f()
t0 = time.perf_counter(); f(); print(which, "Basis + upload s", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); f(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:4200])
