"""Golden vectors for FacetBasis, produced by the REAL reference (scikit-fem
12.0.1 imported read-only from /root/reference) in the build container.

    python tools/gen_golden_facet.py        -> tests/golden/facet_*.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import skfem as fem  # noqa: E402
from skfem.helpers import dot, grad  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


@fem.BilinearForm
def bmass(u, v, w):
    return u * v


@fem.BilinearForm
def nitsche(u, v, w):
    # the boundary terms of Nitsche's method (docs/examples/ex14-style)
    return 1. / (1e-2 * w.h) * u * v - dot(w.n, grad(u)) * v - dot(w.n, grad(v)) * u


@fem.BilinearForm
def robin(u, v, w):
    return (2. + w.x[0] * w['prev']) * u * v


@fem.LinearForm
def flux(v, w):
    return w.x[0] * v + dot(w.n, grad(v)) * w.x[1]


@fem.LinearForm
def coef_load(v, w):
    return dot(w['prev'].grad, w.n) * v


@fem.Functional
def area(w):
    return 1.


@fem.Functional
def divthm(w):
    return w.n[0] * w.x[0]


@fem.BilinearForm
def vtraction(u, v, w):
    return dot(u, w.n) * dot(v, w.n) + 0.5 * dot(u, v)


@fem.BilinearForm
def jump(u, v, w):
    return u * v + dot(grad(u), w.n) * v


def csr(prefix, A):
    return {prefix + "_indptr": A.indptr, prefix + "_indices": A.indices,
            prefix + "_data": A.data, prefix + "_shape": np.array(A.shape)}


def dump(name, m, e, vector=False):
    fb = fem.FacetBasis(m, e)
    out = dict(p=m.p, t=m.t, find=fb.find, tind=fb.tind, X=fb.X, W=fb.W, dx=fb.dx,
               normals=np.asarray(fb.normals), x=np.asarray(fb.global_coordinates()),
               h=np.asarray(fb.mesh_parameters()), element_dofs=fb.element_dofs,
               N=np.int64(fb.N), facets=m.facets, f2t=m.f2t,
               phi=np.array([np.asarray(b[0]) for b in fb.basis]),
               dphi=np.array([b[0].grad for b in fb.basis]))
    if vector:
        out["vtraction_local"] = vtraction.elemental(fb).data
        out.update(csr("vtraction", vtraction.assemble(fb)))
    else:
        cb = fem.CellBasis(m, e)
        prev = np.cos(2. * cb.doflocs[0]) + cb.doflocs[1] ** 2
        out["prev"] = prev
        for nm, form, kw in [("bmass", bmass, {}), ("nitsche", nitsche, {}),
                             ("robin", robin, dict(prev=prev))]:
            out[nm + "_local"] = form.elemental(fb, **kw).data
            out.update(csr(nm, form.assemble(fb, **kw)))
        out["flux_vec"] = flux.assemble(fb)
        out["coef_load_vec"] = coef_load.assemble(fb, prev=prev)
        out["area_elemental"] = area.elemental(fb)
        out["area"] = np.float64(area.assemble(fb))
        out["divthm"] = np.float64(divthm.assemble(fb))
        out["divthm_elemental"] = divthm.elemental(fb)
        # a subset of the boundary through a callable, and interior facets from
        # both sides (jump-type coupling between the two traces)
        sub = fem.FacetBasis(m, e, facets=m.facets_satisfying(lambda x: x[0] < 0.3,
                                                              boundaries_only=True))
        out["sub_find"] = sub.find
        out.update(csr("sub_bmass", bmass.assemble(sub)))
        interior = np.nonzero(m.f2t[1] != -1)[0][::3].astype(np.int32)
        f0 = fem.FacetBasis(m, e, facets=interior, side=0)
        f1 = fem.FacetBasis(m, e, facets=interior, side=1)
        out["interior_find"] = interior
        out["jump_local"] = jump.elemental(f0, f1).data
        out.update(csr("jump", jump.assemble(f0, f1)))
        out["interior_normals"] = np.asarray(f1.normals)
        out["interior_phi1"] = np.array([np.asarray(b[0]) for b in f1.basis])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "facets", len(fb.find), "nqp", fb.W.shape[0])


mt = fem.MeshTri().refined(3)
mt = fem.MeshTri(np.vstack((mt.p[0] + 0.05 * np.sin(5 * mt.p[1]), mt.p[1] + 0.1 * mt.p[0] ** 2)),
                 mt.t)
dump("facet_tri_p1", mt, fem.ElementTriP1())
dump("facet_tri_p2", mt, fem.ElementTriP2())

x = np.linspace(0, 1, 4)
mx = fem.MeshTet.init_tensor(x, np.linspace(0, 1, 3), np.linspace(0, 1, 4))
q = mx.p.copy()
q[0] = mx.p[0] + 0.03 * np.sin(7 * mx.p[1])
q[1] = mx.p[1] + 0.02 * mx.p[2] ** 2
mx = fem.MeshTet(q, mx.t)
dump("facet_tet_p1", mx, fem.ElementTetP1())
dump("facet_tet_p2", mx, fem.ElementTetP2())
dump("facet_tet_vp1", mx, fem.ElementVector(fem.ElementTetP1()), vector=True)
