"""Small warm assembly through the fused P1 kernel for compute-sanitizer runs:

    compute-sanitizer --tool racecheck python tools/sanitizer_case.py
    compute-sanitizer --tool memcheck  python tools/sanitizer_case.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200"))
sys.path.insert(0, ROOT)
import skfem_b200 as fem  # noqa: E402
from skfem_b200.models.poisson import laplace  # noqa: E402
from oracle import skfem_oracle as O  # noqa: E402

x = np.linspace(0, 1, 13)
m = fem.MeshTet.init_tensor(x, x, x)                       # 10 368 tets: 41 tiles, 11 super-tiles
b = fem.Basis(m, fem.ElementTetP1())
laplace.assemble(b)
A = laplace.assemble(b)                                    # fused kernel + combine
b.update_points(m.p * np.array([[1.25], [0.8], [1.1]]))
A2 = laplace.assemble(b)
Ao = O.assemble_bilinear(O.laplace, O.cell_basis(O.mesh_tet_tensor(x, x, x), O.element("tet_p1")))
assert np.array_equal(A.indices, Ao.indices)
np.testing.assert_allclose(A.data, Ao.data, rtol=1e-12, atol=1e-12 * np.abs(Ao.data).max())
print("sanitizer case ok", A.nnz, float(np.abs(A2.data).sum()))
