"""Golden vectors for FacetBasis on oriented facet sets (OrientedBoundary: Mesh.facets_around,
Mesh.facets_satisfying(normal=...); facet_basis.py:84-89), produced by the REAL reference
(scikit-fem 12.0.1, /root/reference).

    python tools/gen_golden_oriented.py        -> tests/golden/facet_oriented_{tri,tet}.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import skfem as fem  # noqa: E402
from skfem.helpers import dot, grad  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


@fem.BilinearForm
def flow(u, v, w):
    return dot(grad(u), w.n) * v + u * v


@fem.Functional
def divthm(w):
    return dot(w.n, w.x)


def dump(name, m, e, normal):
    out = dict(p=m.p, t=m.t)
    # a subdomain and a band of facets strictly inside the domain (on boundary facets the
    # reference indexes f2t[1] == -1, i.e. the last element - nothing worth pinning)
    def interior(x):
        return np.all((x > 0.2) * (x < 0.8), axis=0)
    inside = m.elements_satisfying(lambda x: interior(x) * (x[0] < 0.55))
    out["inside"] = inside
    sets = {"around": m.facets_around(inside), "around_flip": m.facets_around(inside, flip=True),
            "normal": m.facets_satisfying(lambda x: interior(x) * (x[0] > 0.3) * (x[0] < 0.7),
                                          normal=normal)}
    for key, ob in sets.items():
        out[key + "_find"], out[key + "_ori"] = np.asarray(ob), ob.ori
        for side in (0, 1):
            if side == 1 and (m.f2t[1, np.asarray(ob)] == -1).any():
                continue
            fb = fem.FacetBasis(m, e, facets=ob, side=side)
            k = "{}_s{}".format(key, side)
            out[k + "_tind"], out[k + "_tind_normals"] = fb.tind, fb.tind_normals
            out[k + "_normals"], out[k + "_dx"] = np.asarray(fb.normals), fb.dx
            A = flow.assemble(fb)
            out[k + "_indptr"], out[k + "_indices"], out[k + "_data"] = A.indptr, A.indices, A.data
            out[k + "_divthm"] = np.float64(divthm.assemble(fb))
            print(name, k, len(fb.find), "ori sum", int(ob.ori.sum()), "divthm", out[k + "_divthm"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


mt = fem.MeshTri().refined(3)
mt = fem.MeshTri(np.vstack((mt.p[0] + 0.05 * np.sin(5 * mt.p[1]), mt.p[1] + 0.1 * mt.p[0] ** 2)),
                 mt.t)
dump("facet_oriented_tri", mt, fem.ElementTriP2(), np.array([1., 0.2]))
x = np.linspace(0, 1, 4)
mx = fem.MeshTet.init_tensor(x, np.linspace(0, 1, 3), np.linspace(0, 1, 4))
q = mx.p.copy()
q[0] = mx.p[0] + 0.03 * np.sin(7 * mx.p[1])
q[1] = mx.p[1] + 0.02 * mx.p[2] ** 2
dump("facet_oriented_tet", fem.MeshTet(q, mx.t), fem.ElementTetP1(), np.array([1., 0.2, -0.1]))
