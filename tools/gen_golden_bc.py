"""Golden vectors for the device-side boundary conditions (enforce / condense /
solve), produced by the REAL reference (scikit-fem 12.0.1, /root/reference).

    python tools/gen_golden_bc.py        -> tests/golden/bc_*.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import skfem as fem  # noqa: E402
from skfem.models.poisson import laplace, mass, unit_load  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def csr(prefix, A):
    A = A.tocsr()
    return {prefix + "_indptr": A.indptr, prefix + "_indices": A.indices,
            prefix + "_data": A.data, prefix + "_shape": np.array(A.shape)}


def dump(name, m, e):
    basis = fem.Basis(m, e)
    A = laplace.assemble(basis)
    M = mass.assemble(basis)
    b = unit_load.assemble(basis)
    D = basis.get_dofs().all()
    x = np.sin(3. * basis.doflocs[0]) + basis.doflocs[1] ** 2      # inhomogeneous data
    out = dict(p=m.p, t=m.t, D=D, x=x, b=b)
    out.update(csr("A", A))
    # enforce: homogeneous, inhomogeneous with diag, matrix rhs
    Ae, be = fem.enforce(A, b, D=D)
    out.update(csr("enf", Ae)); out["enf_b"] = be
    Ae2, be2 = fem.enforce(A, b, x=x, D=D, diag=2.5)
    out.update(csr("enf2", Ae2)); out["enf2_b"] = be2
    Ae3, Me3 = fem.enforce(A, M, D=D)
    out.update(csr("enf3_M", Me3))
    # condense: homogeneous, inhomogeneous, I given, matrix rhs, no rhs
    AII, bI, xI, I = fem.condense(A, b, D=D)
    out.update(csr("con", AII)); out["con_b"] = bI; out["con_I"] = I
    AII2, bI2, x2, I2 = fem.condense(A, b, x=x, D=D)
    out.update(csr("con2", AII2)); out["con2_b"] = bI2
    AII3, bI3 = fem.condense(A, b, x=x, I=I, expand=False)
    assert (AII3 != AII2).nnz == 0 and np.array_equal(bI3, bI2)
    AII4, MII4, _, _ = fem.condense(A, M, D=D)
    out.update(csr("con4_M", MII4))
    A5 = fem.condense(A, x=x, D=D, expand=False)            # b = zeros_like(x)
    out["con5_b"] = A5[1]
    # solutions (direct solver of the reference)
    out["sol"] = fem.solve(*fem.condense(A, b, D=D))
    out["sol2"] = fem.solve(*fem.condense(A, b, x=x, D=D))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N", basis.N, "nD", len(D))


mt = fem.MeshTri().refined(3)
dump("bc_tri_p1", mt, fem.ElementTriP1())
xs = np.linspace(0, 1, 4)
mx = fem.MeshTet.init_tensor(xs, xs, np.linspace(0, 1, 3))
q = mx.p.copy()
q[0] = mx.p[0] + 0.03 * np.sin(7 * mx.p[1])
dump("bc_tet_p2", fem.MeshTet(q, mx.t), fem.ElementTetP2())


# ---- get_dofs with facet selectors (abstract_basis.py:124-237) -----------------------
def dofs_case(m, e):
    mb = m.with_boundaries({'left': lambda x: np.isclose(x[0], 0.),
                            'top': lambda x: np.isclose(x[1], 1.)})
    basis = fem.Basis(mb, e)
    return dict(all=basis.get_dofs().all(), left=basis.get_dofs('left').all(),
                both=basis.get_dofs({'left', 'top'}).all(),
                fn=basis.get_dofs(lambda x: x[0] > 0.6).all(),
                left_facets=mb.boundaries['left'], top_facets=mb.boundaries['top'])


out = {}
xs3 = np.linspace(0, 1, 4)
for name, m, e in [("tri_p2", fem.MeshTri().refined(2), fem.ElementTriP2()),
                   ("tet_p2", fem.MeshTet.init_tensor(xs3, xs3, xs3), fem.ElementTetP2()),
                   ("tet_vp1", fem.MeshTet.init_tensor(xs3, xs3, xs3),
                    fem.ElementVector(fem.ElementTetP1())),
                   ("hex2", fem.MeshHex.init_tensor(xs3, xs3, xs3), fem.ElementHex2())]:
    for k, v in dofs_case(m, e).items():
        out[name + "_" + k] = v
np.savez_compressed(os.path.join(OUT, "bc_get_dofs.npz"), **out)
print("bc_get_dofs", sorted(out)[:4], "...")
