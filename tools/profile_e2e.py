"""Host-side breakdown of the cold end-to-end call (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import skfem_b200 as fem
from skfem_b200.models.poisson import laplace
from skfem_b200 import form as F

x = np.linspace(0, 1, 101)
m0 = fem.MeshTet.init_tensor(x, x, x)
p = torch.from_numpy(m0.p).pin_memory().numpy(); t = torch.from_numpy(m0.t).pin_memory().numpy()

def T(label, fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"{label:40s} {1e3*dt:8.2f} ms"); return r

for rep in range(2):
    print("--- pass", rep)
    m = T("MeshTet(p,t)", lambda: fem.MeshTet(p, t))
    b = T("Basis(m, TetP1)", lambda: fem.Basis(m, fem.ElementTetP1()))
    d = T("basis._dev() [H2D]", lambda: fem.Basis(m, fem.ElementTetP1())._dev())
    def cold():
        bb = fem.Basis(m, fem.ElementTetP1()); return laplace.assemble_device(bb)
    A = T("Basis + assemble_device (cold)", cold)
    bb = fem.Basis(m, fem.ElementTetP1()); bb._dev()
    loc = T("  _local (generic kernel)", lambda: laplace._local(bb))
    dd = bb._dev()
    pl = T("  build_plan", lambda: F.build_plan(dd["edofs"], dd["edofs"], bb.nelems, (bb.N, bb.N), loc))
    T("to_scipy [D2H]", lambda: A.to_scipy())
    def full():
        mm = fem.MeshTet(p, t); bb = fem.Basis(mm, fem.ElementTetP1()); return laplace.assemble(bb)
    T("full e2e", full)

# per-iteration wall time of the full cold call (detects allocator warm-up / bimodality)
print("--- per-iteration full e2e (ms)")
res = None
ts = []
for i in range(14):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = full()
    torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
print(" ".join("%.1f" % v for v in ts))
# the same after a CUDA-graph capture (torch empties its allocator cache on capture)
g = torch.cuda.CUDAGraph()
bb2 = fem.Basis(m, fem.ElementTetP1()); laplace.assemble_device(bb2); out = torch.empty(laplace.assemble_device(bb2).nnz, dtype=torch.float64, device="cuda")
with torch.cuda.graph(g):
    laplace.assemble_device(bb2, out=out)
ts = []
for i in range(14):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = full()
    torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
print("after graph capture:", " ".join("%.1f" % v for v in ts))
