"""Two cold P1 Laplace assemblies of BASELINE configs[1] through the public API (for an ncu launch
list of the cold path: `ncu --metrics gpu__time_duration.sum --csv python tools/cold_once.py`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import skfem_b200 as fem
from skfem_b200.models.poisson import laplace
from skfem_b200 import form as F

if len(sys.argv) > 1:
    F.set_options(plan_method=sys.argv[1])
x = np.linspace(0, 1, 101)
m = fem.MeshTet.init_tensor(x, x, x)
for rep in range(2):
    torch.cuda.nvtx.range_push("cold%d" % rep)
    A = laplace.assemble(fem.Basis(fem.MeshTet(m.p, m.t), fem.ElementTetP1()))
    torch.cuda.nvtx.range_pop()
print(A.nnz)
