import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scikit-fem_b200")); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import skfem_b200 as fem
from skfem_b200 import _lib
from skfem_b200.models.elasticity import linear_elasticity, lame_parameters
from skfem_b200.models.poisson import laplace
x = np.linspace(0, 1, 40)
m = fem.MeshTet.init_tensor(x, x, x)
for name, elem, form in (("vecP2 elasticity", fem.ElementVector(fem.ElementTetP2()), linear_elasticity(*lame_parameters(1e3, .3))), ("P2 laplace", fem.ElementTetP2(), laplace)):
    b = fem.Basis(m, elem)
    for flag in (0, 8):
        _lib.lib().skb_debug_flags(flag)
        form._local(b); torch.cuda.synchronize()
        t0 = time.perf_counter(); loc = form._local(b); torch.cuda.synchronize(); t1 = time.perf_counter()
        print(name, "flag", flag, "local kernel ms", round(1e3*(t1-t0), 2), "nel", b.nelems)
        del loc
    _lib.lib().skb_debug_flags(0)
    A = form.assemble_device(b); torch.cuda.synchronize()
    t0 = time.perf_counter(); A = form.assemble_device(b); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(name, "warm assemble ms", round(1e3*(t1-t0), 2))
    del A, b
    torch.cuda.empty_cache()
