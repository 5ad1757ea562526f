"""Generate golden input/output vectors by running the REAL reference
(scikit-fem 12.0.1 imported read-only from /root/reference) in the build
container.  The GPU box has no /root/reference, so the vectors are committed
under tests/golden/ together with this script (task brief, section 3).

Each fixture stores the inputs (p, t), the DOF numbering, the quadrature rule,
the element-local data ``Form.elemental(basis).data`` and the assembled CSR /
load vector exactly as the reference produced them.

    python tools/gen_golden.py
"""
import sys
import os
import numpy as np

sys.path.insert(0, "/root/reference")
import skfem as fem  # noqa: E402
from skfem.models.poisson import laplace, mass, unit_load, vector_laplace  # noqa
from skfem.models.elasticity import linear_elasticity, lame_parameters  # noqa
from skfem.helpers import dot, grad  # noqa

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def lin(n):
    return np.linspace(0, 1, n)


def morph(m):
    return m.morphed(lambda p: p[0] + 0.03 * np.sin(7 * p[1]),
                     lambda p: p[1] + 0.02 * p[2] ** 2,
                     None) if hasattr(m, "morphed") else None


def morph_pts(p):
    q = p.copy()
    q[0] = p[0] + 0.03 * np.sin(7 * p[1])
    q[1] = p[1] + 0.02 * p[2] ** 2
    return q


@fem.BilinearForm
def user_aniso(u, v, w):
    # a user-defined (non-library) form: coefficient depends on w.x
    return (1. + w.x[0] * w.x[1]) * dot(grad(u), grad(v)) + 3. * u * v


@fem.LinearForm
def user_load(v, w):
    return np.sin(3. * w.x[0]) * v + w.x[1] * v


def dump(name, m, e, bil=(), lin_=(), store_local=True):
    b = fem.Basis(m, e)
    out = dict(p=m.p, t=m.t, element_dofs=b.element_dofs, N=np.int64(b.N),
               X=b.X, W=b.W, dx=b.dx if b.dx.size < 200000 else b.dx[:64])
    es = e.elem if isinstance(e, fem.ElementVector) else e
    nbs = b.Nbfun // (e.dim if isinstance(e, fem.ElementVector) else 1)
    tabs = [es.lbasis(b.X, i) for i in range(nbs)]
    out["phi"] = np.array([np.broadcast_to(t_[0], b.W.shape) for t_ in tabs])
    out["dphi"] = np.array([t_[1] for t_ in tabs])
    out["x"] = b.global_coordinates() if m.t.shape[1] < 3000 else np.zeros(0)
    for fname, form in bil:
        coo = form.elemental(b)
        A = coo.tocsr()
        assert A.has_canonical_format
        if store_local:
            out[f"{fname}_local"] = coo.data
        out[f"{fname}_indptr"] = A.indptr
        out[f"{fname}_indices"] = A.indices
        out[f"{fname}_data"] = A.data
    for fname, form in lin_:
        out[f"{fname}_vec"] = form.assemble(b)
        if store_local:
            out[f"{fname}_local"] = form.elemental(b).data
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: nel={m.t.shape[1]} N={b.N} "
          + " ".join(f"{f}:nnz={len(out[f + '_data'])}" for f, _ in bil)
          + f"  ({os.path.getsize(path) / 1e3:.0f} kB)")


lam, mu = lame_parameters(1e3, 0.3)
elast = linear_elasticity(lam, mu)

# C1: README ex01 shape
dump("c1_tri_p1_refined4", fem.MeshTri().refined(4), fem.ElementTriP1(),
     bil=[("laplace", laplace), ("mass", mass), ("user_aniso", user_aniso)],
     lin_=[("unit_load", unit_load), ("user_load", user_load)])
# known-answer doctest mesh (assembly/__init__.py:38-46)
dump("tri_p1_two_triangles", fem.MeshTri(), fem.ElementTriP1(),
     bil=[("laplace", laplace), ("mass", mass)], lin_=[("unit_load", unit_load)])
# C2 shape, small
dump("tet_p1_tensor6", fem.MeshTet.init_tensor(lin(7), lin(7), lin(7)),
     fem.ElementTetP1(),
     bil=[("laplace", laplace), ("mass", mass), ("user_aniso", user_aniso)],
     lin_=[("unit_load", unit_load), ("user_load", user_load)])
# non-uniform tensor grid (different spacing per axis)
dump("tet_p1_tensor_nonuniform",
     fem.MeshTet.init_tensor(lin(5) ** 2, lin(4), np.sqrt(lin(6))),
     fem.ElementTetP1(), bil=[("laplace", laplace), ("mass", mass)],
     lin_=[("unit_load", unit_load)])
# unstructured: ball, refined default, morphed
dump("tet_p1_ball2", fem.MeshTet.init_ball(2), fem.ElementTetP1(),
     bil=[("laplace", laplace), ("mass", mass)], lin_=[("unit_load", unit_load)])
dump("tet_p1_refined3", fem.MeshTet().refined(3), fem.ElementTetP1(),
     bil=[("laplace", laplace)], lin_=[("unit_load", unit_load)])
mt = fem.MeshTet.init_tensor(lin(6), lin(6), lin(6))
mm = fem.MeshTet(morph_pts(mt.p), mt.t)
dump("tet_p1_morphed5", mm, fem.ElementTetP1(),
     bil=[("laplace", laplace), ("mass", mass)], lin_=[("unit_load", unit_load)])
# P2 scalar
dump("tet_p2_tensor3", fem.MeshTet.init_tensor(lin(4), lin(4), lin(4)),
     fem.ElementTetP2(), bil=[("laplace", laplace), ("mass", mass)],
     lin_=[("unit_load", unit_load)])
m3 = fem.MeshTet.init_tensor(lin(4), lin(4), lin(4))
dump("tet_p2_morphed3", fem.MeshTet(morph_pts(m3.p), m3.t),
     fem.ElementTetP2(),
     bil=[("laplace", laplace), ("mass", mass), ("user_aniso", user_aniso)],
     lin_=[("unit_load", unit_load), ("user_load", user_load)])
# C3 shape, small: vector P2 elasticity (+ vector P1 variants)
dump("tet_vp2_elasticity2", fem.MeshTet.init_tensor(lin(3), lin(3), lin(3)),
     fem.ElementVector(fem.ElementTetP2()),
     bil=[("elasticity", elast), ("vector_laplace", vector_laplace)])
m2 = fem.MeshTet.init_tensor(lin(3), lin(3), lin(3))
dump("tet_vp2_elasticity_morphed2", fem.MeshTet(morph_pts(m2.p), m2.t),
     fem.ElementVector(fem.ElementTetP2()), bil=[("elasticity", elast)])
dump("tet_vp1_elasticity4", fem.MeshTet.init_tensor(lin(5), lin(5), lin(5)),
     fem.ElementVector(fem.ElementTetP1()),
     bil=[("elasticity", elast), ("vector_laplace", vector_laplace),
          ("mass", fem.BilinearForm(lambda u, v, w: dot(u, v)))])
# C4 shape, small: hexes
dump("hex1_tensor3", fem.MeshHex.init_tensor(lin(4), lin(4), lin(4)),
     fem.ElementHex1(), bil=[("laplace", laplace), ("mass", mass)],
     lin_=[("unit_load", unit_load)])
mh = fem.MeshHex.init_tensor(lin(4), lin(4), lin(4))
dump("hex1_morphed3", fem.MeshHex(morph_pts(mh.p), mh.t), fem.ElementHex1(),
     bil=[("laplace", laplace), ("mass", mass)], lin_=[("unit_load", unit_load)])
dump("hex2_tensor2", fem.MeshHex.init_tensor(lin(3), lin(3), lin(3)),
     fem.ElementHex2(), bil=[("laplace", laplace), ("mass", mass)],
     lin_=[("unit_load", unit_load)], store_local=False)
mh2 = fem.MeshHex.init_tensor(lin(3), lin(3), lin(3))
dump("hex2_morphed2", fem.MeshHex(morph_pts(mh2.p), mh2.t), fem.ElementHex2(),
     bil=[("laplace", laplace), ("mass", mass)], store_local=False)
# reference tables of Hex2 at its default rule (element_hex2.py:11-1210 is
# machine-generated Horner code that cannot be restated independently)
e2 = fem.ElementHex2()
X, W = fem.quadrature.get_quadrature(e2.refdom, 2 * e2.maxdeg)
tab = [e2.lbasis(X, i) for i in range(27)]
np.savez_compressed(os.path.join(OUT, "hex2_tables.npz"), X=X, W=W,
                    phi=np.array([t[0] for t in tab]),
                    dphi=np.array([t[1] for t in tab]))

# ---- the traced surface: TriP2, coefficient fields (interpolate), Functional,
# two different bases, asm() with a bare callable, Form.partial -----------------
mt2 = fem.MeshTri().refined(3)
mt2 = fem.MeshTri(np.vstack((mt2.p[0] + 0.05 * np.sin(5 * mt2.p[1]), mt2.p[1])), mt2.t)
dump("tri_p2_morphed3", mt2, fem.ElementTriP2(),
     bil=[("laplace", laplace), ("mass", mass), ("user_aniso", user_aniso)],
     lin_=[("unit_load", unit_load), ("user_load", user_load)])

mx = fem.MeshTet.init_tensor(lin(5), lin(4), lin(5))
mx = fem.MeshTet(morph_pts(mx.p), mx.t)
b1 = fem.Basis(mx, fem.ElementTetP1())
b2 = fem.Basis(mx, fem.ElementTetP2(), intorder=2)        # same rule as b1 (4 points)
prev = np.cos(2. * b1.doflocs[0]) + b1.doflocs[1] * b1.doflocs[2]


@fem.BilinearForm
def newton_like(u, v, w):
    # coefficient field + its gradient, as in docs/examples/ex10.py
    return (1. + w['prev'] ** 2) * dot(grad(u), grad(v)) + dot(w['prev'].grad, grad(v)) * u


@fem.LinearForm
def residual_like(v, w):
    return dot(w['prev'].grad, grad(v)) + w['prev'] * v * w['scale']


@fem.Functional
def energy(w):
    return 0.5 * dot(w['prev'].grad, w['prev'].grad) + w.x[0] * w['prev']


@fem.BilinearForm
def mixed(u, v, w):
    return u * v + dot(grad(u), grad(v))


def scaled_mass(u, v, w):
    return w['alpha'] * u * v


A_n = newton_like.assemble(b1, prev=prev)
r_n = residual_like.assemble(b1, prev=prev, scale=2.5)
A_mix = mixed.assemble(b2, b1)                        # u in P2, v in P1  -> (N1, N2)
A_asm = fem.asm(scaled_mass, b1, alpha=3.0)
np.savez_compressed(
    os.path.join(OUT, "traced_surface.npz"), p=mx.p, t=mx.t, prev=prev,
    newton_local=newton_like.elemental(b1, prev=prev).data,
    newton_indptr=A_n.indptr, newton_indices=A_n.indices, newton_data=A_n.data,
    residual_vec=r_n, energy=np.float64(energy.assemble(b1, prev=prev)),
    energy_elemental=energy.elemental(b1, prev=prev),
    mixed_indptr=A_mix.indptr, mixed_indices=A_mix.indices, mixed_data=A_mix.data,
    mixed_local=mixed.elemental(b2, b1).data, mixed_shape=np.array(A_mix.shape),
    asm_indptr=A_asm.indptr, asm_indices=A_asm.indices, asm_data=A_asm.data,
    interp_value=np.asarray(b1.interpolate(prev)), interp_grad=b1.interpolate(prev).grad,
    doflocs=b1.doflocs)
print("traced_surface: newton nnz", A_n.nnz, "mixed", A_mix.shape, A_mix.nnz)

# closed-form / known-answer facts
b = fem.Basis(fem.MeshTri().refined(4), fem.ElementTriP1())
A = laplace.assemble(b)
f = unit_load.assemble(b)
x = fem.solve(*fem.condense(A, f, D=b.get_dofs()))
print("ex01 max(x) =", repr(x.max()))
np.savez_compressed(os.path.join(OUT, "ex01_solution.npz"), x=x,
                    D=b.get_dofs().flatten())
