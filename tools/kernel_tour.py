"""Launch every kernel of libskfem_b200.so once at a moderate size through the public API (for
`ncu --set full -k regex:skb`): P1 / P2 / vector-P2 tets, Hex1 / Hex2, linear forms, the traced
path, FacetBasis, boundary conditions + SpMV, the fused warm path and both plan builders."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import skfem_b200 as fem
from skfem_b200 import form as F
from skfem_b200.helpers import dot, grad
from skfem_b200.models.poisson import laplace, mass, unit_load
from skfem_b200.models.elasticity import linear_elasticity, lame_parameters

n = int(sys.argv[1]) if len(sys.argv) > 1 else 41
x = np.linspace(0, 1, n)
mt = fem.MeshTet.init_tensor(x, x, x)
b1 = fem.Basis(mt, fem.ElementTetP1())
A = laplace.assemble_device(b1)            # local_affine, rows plan, csr_reduce
F.set_options(plan_method="sort")
laplace.assemble_device(fem.Basis(mt, fem.ElementTetP1()))   # radix-sort plan kernels
F.set_options(plan_method="rows")
laplace.assemble_device(b1)                # fused plan build + fused kernel + combine
laplace.assemble_device(b1)
unit_load.assemble_device(b1)              # local_linear, vec_reduce
mass.assemble_device(b1)                   # sym kernel; then the fused mass path (MODE 4)
mass.assemble_device(b1)
mass.assemble_device(b1)
xs = np.linspace(0, 1, max((n + 1) // 2, 25))   # >= 65 536 tets: entities / DOF tables on the GPU
ms = fem.MeshTet.init_tensor(xs, xs, xs)
b2 = fem.Basis(ms, fem.ElementTetP2())
laplace.assemble_device(b2)
laplace.assemble_device(b2)                # warm: element-major sym kernel + csr_reduce_em
bv = fem.Basis(ms, fem.ElementVector(fem.ElementTetP2()))
el = linear_elasticity(*lame_parameters(1e3, 0.3))
el.assemble_device(bv)                     # cached vector kernel
el.assemble_device(bv)                     # warm: its element-major variant
xh = np.linspace(0, 1, max((n + 1) // 2, 42))   # >= 65 536 hexes
mh = fem.MeshHex.init_tensor(xh, xh, xh)
laplace.assemble_device(fem.Basis(mh, fem.ElementHex1()))           # local_hex
bh2 = fem.Basis(mh, fem.ElementHex2())
laplace.assemble_device(bh2)                                         # sum-factorised kernel
laplace.assemble_device(bh2)                                         # warm: element-major output
mass.assemble_device(bh2)
F.set_options(hex_sumfact=False)
laplace.assemble_device(fem.Basis(mh, fem.ElementHex2()))            # DMMA Gram kernel
F.set_options(hex_sumfact=True)
# traced form with a coefficient field: tabulate + qp_reduce
k = b1.interpolate(np.ones(b1.N))
fem.BilinearForm(lambda u, v, w: w["k"] * dot(grad(u), grad(v))).assemble_device(b1, k=k)
fem.LinearForm(lambda v, w: w.x[0] * v).assemble_device(b1)
# FacetBasis (boundary mass) and boundary conditions on the device CSR
fb = fem.FacetBasis(mt, fem.ElementTetP1())
fem.BilinearForm(lambda u, v, w: u * v).assemble_device(fb)
rhs = unit_load.assemble_device(b1)
D = b1.get_dofs().all()
Ae, be = fem.enforce(A, rhs, D=D)
Ac, bc, xc, I = fem.condense(A, rhs, D=D)
torch.cuda.synchronize()
print("tour done", A.nnz)
