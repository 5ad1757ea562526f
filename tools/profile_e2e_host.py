"""cProfile of the cold end-to-end call (host side), run on the GPU box."""
import os, sys, cProfile, pstats, io, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import skfem_b200 as fem
from skfem_b200.models.poisson import laplace

x = np.linspace(0, 1, 101)
m0 = fem.MeshTet.init_tensor(x, x, x)
p = torch.from_numpy(m0.p).pin_memory().numpy(); t = torch.from_numpy(m0.t).pin_memory().numpy()

def full():
    mm = fem.MeshTet(p, t); bb = fem.Basis(mm, fem.ElementTetP1()); return laplace.assemble(bb)

for _ in range(5):
    full()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    full()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
