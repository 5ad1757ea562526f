"""``basis.get_dofs(...)`` returns a DofsView like the reference's (abstract_basis.py:124-251,
dofs.py:17-262,536-663): facet / element / node selectors, ``skip``, ``all(name)`` / ``keep`` /
``drop`` and the per-entity dictionaries, against tests/golden/dofs_views.npz written by the real
reference (tools/gen_golden_dofsview.py - the case and selector lists below mirror that file)."""
import numpy as np
import pytest

import skfem_b200 as fem
from cases import load

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


def _basis(M, E, vector):
    m = fem.MeshTri().refined(2) if M == "MeshTri" else getattr(fem, M).init_tensor(XS, XS, XS)
    m = m.with_boundaries({"left": lambda x: np.isclose(x[0], 0.),
                           "top": lambda x: np.isclose(x[1], 1.)})
    e = getattr(fem, E)()
    return fem.Basis(m, fem.ElementVector(e) if vector else e)


@pytest.mark.parametrize("case,M,E,vector", CASES)
def test_dofs_views_match_reference(case, M, E, vector):
    g = load("dofs_views")
    basis = _basis(M, E, vector)
    names = list(dict.fromkeys(basis.elem.dofnames))
    checked = 0
    for sname, sel in selectors(basis.mesh.dim(), vector).items():
        view = basis.get_dofs(**sel)
        prefix = case + "/" + sname
        flat = g[prefix + "/flat"]
        assert view.flatten().dtype == np.int32
        assert np.array_equal(view.flatten(), flat) and np.array_equal(np.asarray(view), flat)
        assert len(view) == len(flat) and list(view) == list(flat)
        for kind in ("nodal", "facet", "edge", "interior"):
            got = getattr(view, kind)
            want = {k.split("/")[-1]: g[k] for k in g.files
                    if k.startswith("{}/{}/".format(prefix, kind))}
            assert set(got) == set(want), (prefix, kind)
            for k in want:
                assert np.array_equal(got[k], want[k]), (prefix, kind, k)
        for nm in names:
            assert np.array_equal(view.all(nm), g["{}/all/{}".format(prefix, nm)])
            assert np.array_equal(view.keep([nm]).flatten(), g["{}/all/{}".format(prefix, nm)])
            assert np.array_equal(view.drop(nm).flatten(), g["{}/drop/{}".format(prefix, nm)])
            checked += 1
    assert checked >= 9
    union = basis.get_dofs("left") | basis.get_dofs("top")
    assert np.array_equal(union.flatten(), g[case + "/union/flat"])
    assert np.array_equal(basis.get_dofs({"left", "top"}).flatten(), g[case + "/union/flat"])


def test_dofs_view_behaves_like_an_index_array():
    basis = _basis("MeshTet", "ElementTetP2", True)
    left = basis.get_dofs("left")
    x = np.zeros(basis.N)
    x[left] = 1.
    assert x.sum() == len(left) and np.array_equal(np.nonzero(x)[0], left.flatten())
    assert np.array_equal(left[:3], left.flatten()[:3])
    assert np.array_equal(basis.complement_dofs(left),
                          np.setdiff1d(np.arange(basis.N), left.flatten()))
    both = basis.get_dofs({"a": "left", "b": "top"})
    assert set(both) == {"a", "b"} and np.array_equal(both["a"].flatten(), left.flatten())
    assert "DofsView" in repr(left) and "nodal" in repr(left)
    with pytest.raises(NotImplementedError):
        basis.mesh.normalize_nodes(3)              # like the reference: arrays, not bare ints
    with pytest.raises(ValueError, match="not found"):
        basis.get_dofs("bottom")
