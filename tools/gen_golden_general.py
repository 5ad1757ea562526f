"""Golden vectors for skfem.models.general (divu with mixed bases) and
skfem.helpers.inv / det / mul / identity in a user form, produced by the REAL
reference (scikit-fem 12.0.1, /root/reference).

    python tools/gen_golden_general.py        -> tests/golden/general_forms.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import skfem as fem  # noqa: E402
from skfem.helpers import det, dot, grad, identity, inv, mul  # noqa: E402
from skfem.models.general import divu  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

x = np.linspace(0, 1, 4)
m = fem.MeshTet.init_tensor(x, np.linspace(0, 1, 3), x)
q = m.p.copy()
q[0] = m.p[0] + 0.03 * np.sin(7 * m.p[1])
q[1] = m.p[1] + 0.02 * m.p[2] ** 2
m = fem.MeshTet(q, m.t)
ub = fem.Basis(m, fem.ElementVector(fem.ElementTetP2()))       # default rule: order 4
pb = fem.Basis(m, fem.ElementTetP1(), intorder=4)
B = divu.assemble(ub, pb)                                        # (N_p, N_u)

disp = 0.05 * np.sin(3. * ub.doflocs[0]) * ub.doflocs[1]


@fem.BilinearForm
def deformed_laplace(u, v, w):
    # Laplacian pulled back through the deformation gradient F = I + grad(disp)
    F = grad(w['disp']) + identity(w['disp'])
    Finv = inv(F)
    return dot(mul(Finv, grad(u)), mul(Finv, grad(v))) * det(F)


A = deformed_laplace.assemble(pb, disp=ub.interpolate(disp))
np.savez_compressed(
    os.path.join(OUT, "general_forms.npz"), p=m.p, t=m.t, disp=disp,
    divu_local=divu.elemental(ub, pb).data, divu_indptr=B.indptr, divu_indices=B.indices,
    divu_data=B.data, divu_shape=np.array(B.shape),
    defo_local=deformed_laplace.elemental(pb, disp=ub.interpolate(disp)).data,
    defo_indptr=A.indptr, defo_indices=A.indices, defo_data=A.data)
print("divu", B.shape, B.nnz, "deformed_laplace", A.shape, A.nnz)
