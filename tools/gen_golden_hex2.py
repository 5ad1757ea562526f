"""Golden vectors for ElementHex2 on a 4^3 morphed hexahedral mesh (64 elements, 729 DOFs),
produced by the REAL reference (scikit-fem 12.0.1, /root/reference): element-local data of
laplace / mass and the assembled CSR, to pin the FP64 tensor-core (Gram) path at rtol 1e-12
and the value-dependent pattern on more than the 8-element meshes of tools/gen_golden.py.

A second fixture, hex2_boxes3, is a non-uniform *tensor* grid of 27 axis-parallel boxes with
power-of-two edge ratios: the geometry where exact cancellations (entries == 0.0, which the
reference drops from the pattern, coo_data.py:35) are most likely to differ between two
summation orders.

    python tools/gen_golden_hex2.py        -> tests/golden/hex2_morphed4.npz, hex2_boxes3.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import skfem as fem  # noqa: E402
from skfem.models.poisson import laplace, mass  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
x = np.linspace(0, 1, 5)
m = fem.MeshHex.init_tensor(x, x, x)
p = m.p.copy()
p[0] = m.p[0] + 0.03 * np.sin(7 * m.p[1])
p[1] = m.p[1] + 0.02 * m.p[2] ** 2
m = fem.MeshHex(p, m.t)
b = fem.Basis(m, fem.ElementHex2())
out = dict(p=m.p, t=m.t, element_dofs=b.element_dofs, N=np.array(b.N))
for name, form in (("laplace", laplace), ("mass", mass)):
    A = form.assemble(b)
    out[name + "_indptr"], out[name + "_indices"], out[name + "_data"] = A.indptr, A.indices, A.data
    out[name + "_local"] = form.elemental(b).data
    print(name, A.shape, A.nnz, "zeros in local data:", int((out[name + "_local"] == 0).sum()))
np.savez_compressed(os.path.join(OUT, "hex2_morphed4.npz"), **out)

m = fem.MeshHex.init_tensor(np.array([0., 1., 3., 4.]), np.array([0., .5, 1., 3.]),
                            np.array([0., .25, 1.25, 1.5]))
b = fem.Basis(m, fem.ElementHex2())
out = dict(p=m.p, t=m.t, element_dofs=b.element_dofs, N=np.array(b.N), X=b.X, W=b.W)
for name, form in (("laplace", laplace), ("mass", mass)):
    A = form.assemble(b)
    out[name + "_indptr"], out[name + "_indices"], out[name + "_data"] = A.indptr, A.indices, A.data
    out[name + "_local"] = form.elemental(b).data
    print("boxes", name, A.shape, A.nnz, "zeros in local data:",
          int((out[name + "_local"] == 0).sum()), "smallest |entry| / largest:",
          np.abs(out[name + "_local"]).min() / np.abs(out[name + "_local"]).max())
np.savez_compressed(os.path.join(OUT, "hex2_boxes3.npz"), **out)
