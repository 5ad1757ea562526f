"""One cold vector-P2 elasticity assembly at BASELINE configs[2] size (for an ncu launch list of
the cold path: local kernel, plan kernels, csr_reduce).  argv[1]: plan method (rows | sort)."""
import sys, time; sys.path.insert(0,"scikit-fem_b200")
import numpy as np, torch, skfem_b200 as fem
from skfem_b200.models.elasticity import linear_elasticity, lame_parameters
from skfem_b200 import form as F
if len(sys.argv) > 1: F.set_options(plan_method=sys.argv[1])
x=np.linspace(0,1,70); m=fem.MeshTet.init_tensor(x,x,x)
b=fem.Basis(m, fem.ElementVector(fem.ElementTetP2()))
form=linear_elasticity(*lame_parameters(1e3,0.3))
torch.cuda.synchronize(); t0=time.perf_counter(); A=form.assemble_device(b); torch.cuda.synchronize()
print("c3 cold assemble_device s", time.perf_counter()-t0, A.nnz)
