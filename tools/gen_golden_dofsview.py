"""Golden vectors for ``basis.get_dofs(...)`` views (abstract_basis.py:124-251, dofs.py:17-262,
536-663) from the real reference: element / node selectors, ``skip``, ``keep`` / ``drop`` /
``all(name)`` and the per-entity dictionaries.  Build container only:

    PYTHONPATH=/root/reference python tools/gen_golden_dofsview.py

The case list is shared with tests/test_host_api.py (``dofsview_cases``): the key of every
stored array is ``<case>/<selector>/<what>``.
"""
import os
import warnings

import numpy as np
import skfem

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "dofs_views.npz")
XS = np.linspace(0, 1, 4)
CASES = [("tri_p2", "MeshTri", "ElementTriP2", False), ("tet_vp2", "MeshTet", "ElementTetP2", True),
         ("hex2", "MeshHex", "ElementHex2", False), ("tet_p1", "MeshTet", "ElementTetP1", False)]


def low(x):
    return x[0] < 0.4


def selectors(dim, vector):
    return {"boundary": {}, "left": dict(facets="left"), "fn": dict(facets=low),
            "elements_fn": dict(elements=low), "elements_list": dict(elements=[0, 3]),
            "nodes_fn": dict(nodes=low), "nodes_point": dict(nodes=(0.,) * dim),
            "nodes_array": dict(nodes=np.array([0, 5])),
            "left_skip": dict(facets="left", skip=["u^1"] if vector else ["u"])}


def mesh_for(fem, M):
    m = fem.MeshTri().refined(2) if M == "MeshTri" else getattr(fem, M).init_tensor(XS, XS, XS)
    return m.with_boundaries({"left": lambda x: np.isclose(x[0], 0.),
                              "top": lambda x: np.isclose(x[1], 1.)})


def describe(view, names, out, prefix):
    out[prefix + "/flat"] = view.flatten()
    for kind in ("nodal", "facet", "edge", "interior"):
        for k, v in getattr(view, kind).items():
            out["{}/{}/{}".format(prefix, kind, k)] = v
    for nm in names:
        out["{}/all/{}".format(prefix, nm)] = view.all(nm)
        out["{}/drop/{}".format(prefix, nm)] = view.drop(nm).flatten()


def main():
    warnings.simplefilter("ignore")
    out = {}
    for case, M, E, vector in CASES:
        m = mesh_for(skfem, M)
        e = getattr(skfem, E)()
        e = skfem.ElementVector(e) if vector else e
        basis = skfem.Basis(m, e)
        names = list(dict.fromkeys(e.dofnames))
        for sname, sel in selectors(m.dim(), vector).items():
            describe(basis.get_dofs(**sel), names, out, case + "/" + sname)
        out[case + "/union/flat"] = (basis.get_dofs("left") | basis.get_dofs("top")).flatten()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
