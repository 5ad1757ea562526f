"""Golden vectors for skfem.models.general.curluv / rot / vrot and skfem.helpers.cross in a
user form, produced by the REAL reference (scikit-fem 12.0.1, /root/reference).

    python tools/gen_golden_curl.py        -> tests/golden/general_curl.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import skfem as fem  # noqa: E402
from skfem.helpers import cross, dot  # noqa: E402
from skfem.models.general import curluv, rot, vrot  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

x = np.linspace(0, 1, 4)
m = fem.MeshTet.init_tensor(x, np.linspace(0, 1, 3), x)
q = m.p.copy()
q[0] = m.p[0] + 0.03 * np.sin(7 * m.p[1])
q[1] = m.p[1] + 0.02 * m.p[2] ** 2
m = fem.MeshTet(q, m.t)
vb = fem.Basis(m, fem.ElementVector(fem.ElementTetP1()))
wdofs = 0.3 * np.cos(2. * vb.doflocs[0]) + vb.doflocs[1] * vb.doflocs[2]
wf = vb.interpolate(wdofs)
A = curluv.assemble(vb)
r = rot.assemble(vb, w=wf)
v = vrot.assemble(vb, w=wf)


@fem.BilinearForm
def cross_form(u, v, w):
    return dot(cross(u, w['w']), v)


C = cross_form.assemble(vb, w=wf)
# 2-D: scalar curl of a vector field, vector curl of a scalar field
m2 = fem.MeshTri().refined(2)
p2 = m2.p.copy()
p2[0] = p2[0] + 0.05 * np.sin(3 * p2[1])
m2 = fem.MeshTri(p2, m2.t)
vb2 = fem.Basis(m2, fem.ElementVector(fem.ElementTriP1()))
sb2 = fem.Basis(m2, fem.ElementTriP1())
A2 = curluv.assemble(vb2, sb2)          # (curl u, v): u vector, v scalar
np.savez_compressed(
    os.path.join(OUT, "general_curl.npz"), p=m.p, t=m.t, wdofs=wdofs,
    curluv_local=curluv.elemental(vb).data, curluv_indptr=A.indptr, curluv_indices=A.indices,
    curluv_data=A.data, rot_vec=r, vrot_vec=v,
    cross_local=cross_form.elemental(vb, w=wf).data, cross_indptr=C.indptr,
    cross_indices=C.indices, cross_data=C.data,
    p2=m2.p, t2=m2.t, curluv2_local=curluv.elemental(vb2, sb2).data, curluv2_indptr=A2.indptr,
    curluv2_indices=A2.indices, curluv2_data=A2.data, curluv2_shape=np.array(A2.shape))
print("curluv", A.shape, A.nnz, "cross", C.nnz, "2d", A2.shape, A2.nnz)
