"""CPU tests of the host-side mirror: mesh constructors, topology, DOF
numbering, quadrature and reference tables must be bit-identical to the
reference's (golden vectors), and the C-ABI library must load and export every
symbol the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

import skfem_b200 as fem
from cases import CASES, load
from product import mesh_from, element_from

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", list(CASES))
def test_dofs_quadrature_tables_match_reference(name):
    refdom, ename, vector, _, _, _ = CASES[name]
    g = load(name)
    m = mesh_from(g, refdom)
    b = fem.Basis(m, element_from(ename, vector))
    assert b.element_dofs.dtype == np.int32
    assert np.array_equal(b.element_dofs, g["element_dofs"])
    assert b.N == int(g["N"])
    assert np.array_equal(b.X, g["X"]) and np.array_equal(b.W, g["W"])
    assert np.array_equal(b._phi, g["phi"])
    assert np.array_equal(b._dphi, g["dphi"])


def test_mesh_constructors_match_reference():
    g = load("c1_tri_p1_refined4")
    m = fem.MeshTri().refined(4)
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])
    assert m.t.dtype == np.int32 and m.p.dtype == np.float64
    lin = np.linspace(0, 1, 7)
    g = load("tet_p1_tensor6")
    m = fem.MeshTet.init_tensor(lin, lin, lin)
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])
    g = load("tet_p1_tensor_nonuniform")
    lin = np.linspace
    m = fem.MeshTet.init_tensor(lin(0, 1, 5) ** 2, lin(0, 1, 4), np.sqrt(lin(0, 1, 6)))
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])
    g = load("hex1_tensor3")
    m = fem.MeshHex.init_tensor(*(3 * (np.linspace(0, 1, 4),)))
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])
    g = load("tet_p1_refined3")                 # MeshTet().refined(3), SURVEY 8d parity extra
    m = fem.MeshTet().refined(3)
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])
    assert m.t.dtype == np.int32
    g = load("tri_p1_two_triangles")
    m = fem.MeshTri()
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])


def test_hex2_tables_match_the_reference():
    """At the default rule ElementHex2.tabulate returns the reference's own numbers (shipped
    table) bit for bit; the tensor-product evaluation used at any other points
    agrees with the reference's generated Horner forms to a few ulp."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "hex2_tables.npz"))
    e = fem.ElementHex2()
    X, W = fem.get_quadrature(e.refdom, 2 * e.maxdeg)
    assert np.array_equal(X, g["X"]) and np.array_equal(W, g["W"])
    phi, dphi = e.tabulate(X)
    assert np.array_equal(phi, g["phi"]) and np.array_equal(dphi, g["dphi"])
    phi_t, dphi_t = fem.element.Element.tabulate(e, X)              # tensor-product fallback
    np.testing.assert_allclose(phi_t, g["phi"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(dphi_t, g["dphi"], rtol=0, atol=2e-14)
    Xo = X[:, ::7] * 0.99                                           # not the shipped points
    po, _ = e.tabulate(Xo)
    np.testing.assert_allclose(po.sum(axis=0), 1.0, rtol=0, atol=1e-14)   # partition of unity
    b = fem.Basis(fem.MeshHex.init_tensor(*(3 * (np.linspace(0, 1, 3),))), e)
    gg = load("hex2_tensor2")
    assert np.array_equal(b.element_dofs, gg["element_dofs"]) and b.N == 125


def test_incompatible_mesh_and_element():
    with pytest.raises(ValueError, match="Incompatible Mesh and Element."):
        fem.Basis(fem.MeshTri(), fem.ElementTetP1())


def test_quadrature_unknown_order():
    with pytest.raises(NotImplementedError):
        fem.get_quadrature(fem.ElementTetP1().refdom, 50)


def test_boundary_dofs():
    m = fem.MeshTri().refined(2)
    b = fem.Basis(m, fem.ElementTriP1())
    D = b.get_dofs()
    on_bnd = np.nonzero((m.p[0] == 0) | (m.p[0] == 1) | (m.p[1] == 0) | (m.p[1] == 1))[0]
    assert np.array_equal(np.sort(D), on_bnd)


@pytest.mark.parametrize("name,M,E", [
    ("facet_tri_p1", "MeshTri", "ElementTriP1"), ("facet_tri_p2", "MeshTri", "ElementTriP2"),
    ("facet_tet_p1", "MeshTet", "ElementTetP1"), ("facet_tet_p2", "MeshTet", "ElementTetP2")])
def test_facet_basis_host_side_matches_reference(name, M, E):
    """facets / f2t / find / tind / facet rule / element_dofs of FacetBasis
    (no device work in the constructor)."""
    g = load(name)
    m = getattr(fem, M)(g["p"], g["t"])
    fb = fem.FacetBasis(m, getattr(fem, E)())
    assert np.array_equal(m.facets, g["facets"]) and np.array_equal(m.f2t, g["f2t"])
    assert np.array_equal(fb.find, g["find"]) and np.array_equal(fb.tind, g["tind"])
    assert np.array_equal(fb.X, g["X"]) and np.array_equal(fb.W, g["W"])
    assert np.array_equal(fb.element_dofs, g["element_dofs"]) and fb.N == int(g["N"])
    sub = m.facets_satisfying(lambda x: x[0] < 0.3, boundaries_only=True)
    assert np.array_equal(sub, g["sub_find"])
    assert np.array_equal(m.normalize_facets([sub[:3], int(sub[-1])]),
                          np.unique(np.r_[sub[:3], sub[-1]]))
    with pytest.raises(ValueError, match="not found"):
        m.normalize_facets("left")


@pytest.mark.parametrize("name,M,E,args", [
    ("tri_p2", "MeshTri", "ElementTriP2", None), ("tet_p2", "MeshTet", "ElementTetP2", 3),
    ("tet_vp1", "MeshTet", "ElementTetP1", 3), ("hex2", "MeshHex", "ElementHex2", 3)])
def test_get_dofs_facet_selectors_match_reference(name, M, E, args):
    """get_dofs(): whole boundary, named boundaries (with_boundaries), a set of
    names, a test on facet midpoints - against basis.get_dofs(...).all() of the
    reference (tests/golden/bc_get_dofs.npz, tools/gen_golden_bc.py)."""
    g = load("bc_get_dofs")
    if args is None:
        m = fem.MeshTri().refined(2)
    else:
        xs = np.linspace(0, 1, 4)
        m = getattr(fem, M).init_tensor(xs, xs, xs)
    m = m.with_boundaries({'left': lambda x: np.isclose(x[0], 0.),
                           'top': lambda x: np.isclose(x[1], 1.)})
    e = getattr(fem, E)()
    if name == "tet_vp1":
        e = fem.ElementVector(e)
    basis = fem.Basis(m, e)
    assert np.array_equal(m.boundaries['left'], g[name + "_left_facets"])
    assert np.array_equal(m.boundaries['top'], g[name + "_top_facets"])
    assert np.array_equal(basis.get_dofs(), g[name + "_all"])
    assert np.array_equal(basis.get_dofs('left'), g[name + "_left"])
    assert np.array_equal(basis.get_dofs({'left', 'top'}), g[name + "_both"])
    assert np.array_equal(basis.get_dofs(lambda x: x[0] > 0.6), g[name + "_fn"])
    with pytest.raises(ValueError, match="not found"):
        basis.get_dofs('bottom')


def test_mesh_and_basis_are_freed_without_the_cyclic_gc():
    """A mesh owns the device copies of p and t: it (and a basis on it) must die by
    reference counting, or re-assembly loops that build new meshes keep ~100 MB per
    iteration alive until the cyclic GC runs (seen as e2e spikes in bench.py)."""
    import gc
    import weakref
    gc.collect()
    gc.disable()
    try:
        m = fem.MeshTet.init_tensor(*(3 * (np.linspace(0, 1, 4),)))
        b = fem.Basis(m, fem.ElementTetP2())
        fb = fem.FacetBasis(m, fem.ElementTetP2())
        b.get_dofs(), m.f2t, b.mapping.mesh, fb.find
        refs = [weakref.ref(o) for o in (m, b, fb)]
        del m, b, fb
        assert all(r() is None for r in refs)
    finally:
        gc.enable()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "skfem_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(skb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from skfem_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 9
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in _lib.lib().skb_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from skfem_b200.models.poisson import laplace
    b = fem.Basis(fem.MeshTri(), fem.ElementTriP1())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        laplace.assemble(b)


def test_uniform_refinement_of_tets_and_hexes_matches_reference():
    """MeshTet / MeshHex .refined against fixtures written by the real reference
    (tools/gen_golden_mesh.py): same vertices, same children, same order, int32."""
    g = load("mesh_refined")
    x, y, z = g["x"], g["y"], g["z"]
    cases = {
        "tet_default_r2": fem.MeshTet().refined(2),
        "tet_tensor_r1": fem.MeshTet.init_tensor(x, y, z).refined(),
        "hex_default_r2": fem.MeshHex().refined(2),
        "hex_tensor_r1": fem.MeshHex.init_tensor(x, y, z).refined(1),
    }
    for name, m in cases.items():
        assert np.array_equal(m.p, g[name + "_p"]), name
        assert np.array_equal(m.t, g[name + "_t"]) and m.t.dtype == np.int32, name
    # a refined mesh is a full citizen: entities and a quadratic basis can be built on it
    m = cases["hex_tensor_r1"]
    assert m.facets.shape[0] == 4 and m.t2e.shape == (12, m.t.shape[1])
    b = fem.Basis(cases["tet_tensor_r1"], fem.ElementTetP2())
    assert b.N == cases["tet_tensor_r1"].p.shape[1] + cases["tet_tensor_r1"].edges.shape[1]
    with pytest.raises(NotImplementedError):
        type("M", (fem.mesh.Mesh,), {})._uniform(m)


def test_mesh_selections_parameters_and_ball_match_reference():
    """init_ball, params, node / element selections, named subdomains and Basis(elements=name)
    against fixtures from the real reference (tools/gen_golden_mesh.py, tools/gen_golden.py)."""
    g = load("tet_p1_ball2")
    m = fem.MeshTet.init_ball(2)
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])
    p0 = fem.MeshTet.init_ball(0).p                   # the octahedron: no negative zeros
    assert not np.signbit(p0[p0 == 0]).any()
    g = load("mesh_refined")
    x, y, z = g["x"], g["y"], g["z"]
    low = lambda q: q[0] < 0.5                                   # noqa: E731
    meshes = {"tet": (fem.MeshTet.init_tensor(x, y, z), fem.ElementTetP2),
              "tri": (fem.MeshTri().refined(3), fem.ElementTriP2),
              "hex": (fem.MeshHex.init_tensor(x, y, z), fem.ElementHex2)}
    for name, (m, E) in meshes.items():
        assert np.array_equal(m.params(), g[name + "_params"]), name
        assert m.param() == g[name + "_params"].max()
        assert np.array_equal(m.interior_nodes(), g[name + "_interior_nodes"])
        assert np.array_equal(m.nodes_satisfying(low), g[name + "_nodes_low"])
        assert np.array_equal(m.nodes_satisfying(low, boundaries_only=True),
                              g[name + "_bnodes_low"])
        assert np.array_equal(m.elements_satisfying(low), g[name + "_elements_low"])
        ms = m.with_subdomains({"low": low, "pick": np.array([0, 2])})
        assert m.subdomains is None and set(ms.subdomains) == {"low", "pick"}
        assert np.array_equal(ms.normalize_elements(["low", "pick"]), g[name + "_norm_names"])
        assert np.array_equal(ms.normalize_elements([4, 1, 1]), g[name + "_norm_list"])
        assert np.array_equal(ms.normalize_elements(True), np.arange(m.t.shape[1]))
        b = fem.Basis(ms, E(), elements="low")
        assert np.array_equal(b.tind, g[name + "_sub_tind"])
        assert b.nelems == int(g[name + "_sub_nelems"])
        named = ms.with_boundaries({"left": lambda q: q[0] == 0.})
        assert set(named.subdomains) == {"low", "pick"} and "left" in named.boundaries
        with pytest.raises(ValueError, match="Subdomain 'top' not found."):
            ms.normalize_elements("top")


def test_cell_basis_conveniences():
    """with_elements / boundary / quadrature / zero_w (cell_basis.py:266-308,
    abstract_basis.py:378-382): thin host-side constructors over tested primitives."""
    m = fem.MeshTet.init_tensor(*(3 * (np.linspace(0, 1, 4),)))
    b = fem.Basis(m, fem.ElementTetP2())
    fb = b.boundary()
    ref = fem.FacetBasis(m, fem.ElementTetP2())
    assert isinstance(fb, fem.FacetBasis) and np.array_equal(fb.find, ref.find)
    assert np.array_equal(fb.X, ref.X) and fb.mapping is b.mapping
    left = b.boundary(lambda x: x[0] == 0., intorder=2)
    assert np.array_equal(left.find, m.facets_satisfying(lambda x: x[0] == 0.))
    assert np.array_equal(left.W, fem.FacetBasis(m, fem.ElementTetP2(), intorder=2).W)
    sub = b.with_elements(lambda x: x[0] < 0.5)
    assert np.array_equal(sub.tind, m.elements_satisfying(lambda x: x[0] < 0.5))
    assert sub.quadrature[0] is not None and np.array_equal(sub.X, b.X)
    assert np.array_equal(sub.element_dofs, b.element_dofs[:, sub.tind])
    assert b.zero_w().shape == (m.t.shape[1], len(b.W)) and sub.zero_w().shape[0] == sub.nelems
    with pytest.raises(NotImplementedError, match="Boundary of subdomain"):
        sub.boundary()


def test_product_never_touches_the_oracle_or_the_reference():
    """The oracle (and the reference copy under oracle/_ref) is test infrastructure: no file
    of the product package may import, open or mention it."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scikit-fem_b200")
    pat = re.compile(r"\boracle\b|/root/reference|import\s+skfem\b|from\s+skfem\b")
    hits = []
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(d, f), errors="replace") as fh:
                    for n, line in enumerate(fh, 1):
                        if pat.search(line):
                            hits.append("{}:{}: {}".format(f, n, line.strip()))
    assert not hits, hits


def test_element_major_request_falls_back_when_no_kernel_writes_it(monkeypatch):
    """``skb_local_bilinear_em`` answers SKB_EINVAL for spaces without an element-major kernel
    (hexahedra with the scalar kernel, non-default rules): the warm path must then use the
    reference layout, and ask only once per (basis, form)."""
    from skfem_b200 import _lib, form as F
    from skfem_b200.models.poisson import laplace
    b = fem.Basis(fem.MeshHex(), fem.ElementHex1())
    calls = []

    class FakeLib:
        def skb_local_bilinear_em(self, *args):
            calls.append(args)
            return _lib.SKB_EINVAL

    monkeypatch.setattr(F._lib, "lib", lambda: FakeLib())
    monkeypatch.setattr(F, "_stream", lambda: None)
    monkeypatch.setattr(type(b), "_dev",
                        lambda self, device=None: {"device": "cpu", "space": _lib.SkbSpace()})
    assert laplace._local_element_major(b) is None
    assert laplace._local_element_major(b) is None
    assert len(calls) == 1
    F.set_options(element_major=False)
    try:
        assert laplace._local_element_major(fem.Basis(fem.MeshHex(), fem.ElementHex1())) is None
        assert len(calls) == 1
    finally:
        F.set_options(element_major=True)
