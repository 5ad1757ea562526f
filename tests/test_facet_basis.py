"""FacetBasis: boundary / interior-facet quadrature.

Golden vectors tests/golden/facet_*.npz come from the real reference
(tools/gen_golden_facet.py).  The CPU test pins the oracle's restatement
bit-for-bit; the GPU test checks the product (csrc/skb_facet.cu + the traced
path) against the same vectors."""
import numpy as np
import pytest

from cases import load, mesh_of

SCALAR = [("facet_tri_p1", "tri", "tri_p1"), ("facet_tri_p2", "tri", "tri_p2"),
          ("facet_tet_p1", "tet", "tet_p1"), ("facet_tet_p2", "tet", "tet_p2")]


def forms(dot, grad):
    def bmass(u, v, w):
        return u * v

    def nitsche(u, v, w):
        return 1. / (1e-2 * w.h) * u * v - dot(w.n, grad(u)) * v - dot(w.n, grad(v)) * u

    def robin(u, v, w):
        return (2. + w.x[0] * w['prev']) * u * v

    def flux(v, w):
        return w.x[0] * v + dot(w.n, grad(v)) * w.x[1]

    def coef_load(v, w):
        return dot(w['prev'].grad, w.n) * v

    def area(w):
        return 1.

    def divthm(w):
        return w.n[0] * w.x[0]

    def vtraction(u, v, w):
        return dot(u, w.n) * dot(v, w.n) + 0.5 * dot(u, v)

    def jump(u, v, w):
        return u * v + dot(grad(u), w.n) * v
    return dict(bmass=bmass, nitsche=nitsche, robin=robin, flux=flux, coef_load=coef_load,
                area=area, divthm=divthm, vtraction=vtraction, jump=jump)


def check_csr(A, g, prefix, exact):
    assert A.shape == tuple(g[prefix + "_shape"])
    assert np.array_equal(A.indptr, g[prefix + "_indptr"])
    assert np.array_equal(A.indices, g[prefix + "_indices"])
    ref = g[prefix + "_data"]
    if exact:
        assert np.array_equal(A.data, ref)
    else:
        np.testing.assert_allclose(A.data, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


# ---------------------------------------------------------------- oracle (CPU)
@pytest.mark.parametrize("name,refdom,ename", SCALAR)
def test_oracle_facet_scalar(name, refdom, ename):
    from oracle import skfem_oracle as O
    g = load(name)
    m = mesh_of(g, refdom)
    F = forms(O.dot, O.grad)
    fb = O.facet_basis(m, O.element(ename))
    assert np.array_equal(O.facets_of(m)[0], g["facets"])
    assert np.array_equal(O.f2t_of(m), g["f2t"])
    assert np.array_equal(fb.find, g["find"]) and np.array_equal(fb.tind, g["tind"])
    assert np.array_equal(fb.X, g["X"]) and np.array_equal(fb.W, g["W"])
    assert np.array_equal(fb.dx, g["dx"])
    assert np.array_equal(fb.extra["n"], g["normals"])
    assert np.array_equal(np.array(fb.x), g["x"]) and np.array_equal(np.array(fb.h), g["h"])
    assert np.array_equal(fb.element_dofs, g["element_dofs"])
    assert np.array_equal(np.array([np.array(b) for b in fb.basis]), g["phi"])
    assert np.array_equal(np.array([b.grad for b in fb.basis]), g["dphi"])
    prev = g["prev"]
    for nm, kw in [("bmass", {}), ("nitsche", {}), ("robin", dict(prev=prev))]:
        idx, data, shape = O.bilinear_coo(F[nm], fb, **kw)
        assert np.array_equal(data, g[nm + "_local"]), nm
        check_csr(O.coo_to_csr(idx, data, shape), g, nm, exact=True)
    assert np.array_equal(O.assemble_linear(F["flux"], fb), g["flux_vec"])
    assert np.array_equal(O.assemble_linear(F["coef_load"], fb, prev=prev), g["coef_load_vec"])
    assert np.array_equal(O.functional_elemental(F["area"], fb), g["area_elemental"])
    assert np.array_equal(O.functional_elemental(F["divthm"], fb), g["divthm_elemental"])
    sub = O.facet_basis(m, O.element(ename), facets=g["sub_find"])
    check_csr(O.assemble_bilinear(F["bmass"], sub), g, "sub_bmass", exact=True)
    f0 = O.facet_basis(m, O.element(ename), facets=g["interior_find"], side=0)
    f1 = O.facet_basis(m, O.element(ename), facets=g["interior_find"], side=1)
    assert np.array_equal(f1.extra["n"], g["interior_normals"])
    idx, data, shape = O.bilinear_coo(F["jump"], f0, vbasis=f1)
    assert np.array_equal(data, g["jump_local"])
    check_csr(O.coo_to_csr(idx, data, shape), g, "jump", exact=True)


def test_oracle_facet_vector():
    from oracle import skfem_oracle as O
    g = load("facet_tet_vp1")
    m = mesh_of(g, "tet")
    fb = O.facet_basis(m, O.element("tet_p1", vector=True))
    assert np.array_equal(fb.element_dofs, g["element_dofs"])
    assert np.array_equal(np.array([np.array(b) for b in fb.basis]), g["phi"])
    idx, data, shape = O.bilinear_coo(forms(O.dot, O.grad)["vtraction"], fb)
    assert np.array_equal(data, g["vtraction_local"])
    check_csr(O.coo_to_csr(idx, data, shape), g, "vtraction", exact=True)


def test_oracle_facet_known_answers():
    """Boundary measure and the divergence theorem on the unit square / cube."""
    from oracle import skfem_oracle as O
    F = forms(O.dot, O.grad)
    m = O.refine_tri(O.mesh_tri_default(), 3)
    fb = O.facet_basis(m, O.element("tri_p1"))
    assert abs(O.functional_elemental(F["area"], fb).sum() - 4.0) < 1e-13
    assert abs(O.functional_elemental(F["divthm"], fb).sum() - 1.0) < 1e-13
    x = np.linspace(0, 1, 4)
    m3 = O.mesh_tet_tensor(x, x, x)
    fb3 = O.facet_basis(m3, O.element("tet_p1"))
    assert abs(O.functional_elemental(F["area"], fb3).sum() - 6.0) < 1e-13
    A = O.assemble_bilinear(F["bmass"], fb3)
    assert abs(A.sum() - 6.0) < 1e-13           # 1^T B 1 = |boundary|


# --------------------------------------------------------------- product (GPU)
def _elem(fem, ename):
    return {"tri_p1": fem.ElementTriP1, "tri_p2": fem.ElementTriP2,
            "tet_p1": fem.ElementTetP1, "tet_p2": fem.ElementTetP2}[ename]()


@pytest.mark.gpu
@pytest.mark.parametrize("name,refdom,ename", SCALAR)
def test_gpu_facet_scalar(name, refdom, ename):
    import skfem_b200 as fem
    from skfem_b200.helpers import dot, grad
    g = load(name)
    m = (fem.MeshTri if refdom == "tri" else fem.MeshTet)(g["p"], g["t"])
    e = _elem(fem, ename)
    F = forms(dot, grad)
    fb = fem.FacetBasis(m, e)
    assert np.array_equal(m.facets, g["facets"]) and np.array_equal(m.f2t, g["f2t"])
    assert np.array_equal(fb.find, g["find"]) and np.array_equal(fb.tind, g["tind"])
    assert np.array_equal(fb.X, g["X"]) and np.array_equal(fb.W, g["W"])
    assert np.array_equal(fb.dx, g["dx"])
    assert np.array_equal(fb.normals.numpy(), g["normals"])
    w = fb.default_parameters()
    assert np.array_equal(w["x"].numpy(), g["x"]) and np.array_equal(w["h"].numpy(), g["h"])
    assert np.array_equal(fb.element_dofs, g["element_dofs"])
    assert np.array_equal(np.array([b[0].numpy() for b in fb.basis]), g["phi"])
    assert np.array_equal(np.array([b[0].grad.numpy() for b in fb.basis]), g["dphi"])
    prev = g["prev"]
    for nm, kw in [("bmass", {}), ("nitsche", {}), ("robin", dict(prev=prev))]:
        form = fem.BilinearForm(F[nm])
        assert np.array_equal(form.elemental(fb, **kw).data, g[nm + "_local"]), nm
        check_csr(form.assemble(fb, **kw), g, nm, exact=False)
    assert np.array_equal(fem.LinearForm(F["flux"]).assemble(fb), g["flux_vec"])
    assert np.array_equal(fem.LinearForm(F["coef_load"]).assemble(fb, prev=prev),
                          g["coef_load_vec"])
    assert np.array_equal(fem.Functional(F["area"]).elemental(fb), g["area_elemental"])
    assert np.array_equal(fem.Functional(F["divthm"]).elemental(fb), g["divthm_elemental"])
    np.testing.assert_allclose(fem.Functional(F["area"]).assemble(fb), float(g["area"]),
                               rtol=1e-14)
    sub = fem.FacetBasis(m, e, facets=m.facets_satisfying(lambda x: x[0] < 0.3,
                                                          boundaries_only=True))
    assert np.array_equal(sub.find, g["sub_find"])
    check_csr(fem.BilinearForm(F["bmass"]).assemble(sub), g, "sub_bmass", exact=False)
    f0 = fem.FacetBasis(m, e, facets=g["interior_find"], side=0)
    f1 = fem.FacetBasis(m, e, facets=g["interior_find"], side=1)
    assert np.array_equal(f1.normals.numpy(), g["interior_normals"])
    assert np.array_equal(np.array([b[0].numpy() for b in f1.basis]), g["interior_phi1"])
    jump = fem.BilinearForm(F["jump"])
    assert np.array_equal(jump.elemental(f0, f1).data, g["jump_local"])
    check_csr(jump.assemble(f0, f1), g, "jump", exact=False)


@pytest.mark.gpu
def test_gpu_facet_vector():
    import skfem_b200 as fem
    from skfem_b200.helpers import dot, grad
    g = load("facet_tet_vp1")
    m = fem.MeshTet(g["p"], g["t"])
    fb = fem.FacetBasis(m, fem.ElementVector(fem.ElementTetP1()))
    assert np.array_equal(fb.element_dofs, g["element_dofs"])
    assert np.array_equal(np.array([b[0].numpy() for b in fb.basis]), g["phi"])
    assert np.array_equal(np.array([b[0].grad.numpy() for b in fb.basis]), g["dphi"])
    form = fem.BilinearForm(forms(dot, grad)["vtraction"])
    assert np.array_equal(form.elemental(fb).data, g["vtraction_local"])
    check_csr(form.assemble(fb), g, "vtraction", exact=False)


@pytest.mark.gpu
def test_gpu_facet_known_answers_and_errors():
    import skfem_b200 as fem
    from skfem_b200.helpers import dot, grad
    from skfem_b200.models.poisson import laplace, unit_load
    F = forms(dot, grad)
    x = np.linspace(0, 1, 9)
    m = fem.MeshTet.init_tensor(x, x, x)
    fb = fem.FacetBasis(m, fem.ElementTetP1())
    assert abs(fem.Functional(F["area"]).assemble(fb) - 6.0) < 1e-12
    assert abs(fem.Functional(F["divthm"]).assemble(fb) - 1.0) < 1e-12
    B = fem.BilinearForm(F["bmass"]).assemble(fb)
    assert abs(B.sum() - 6.0) < 1e-12
    # library forms on a FacetBasis go through the traced path as well
    L = laplace.assemble(fb)
    assert L.shape == (fb.N, fb.N) and abs(L.sum()) < 1e-10
    assert abs(unit_load.assemble(fb).sum() - 6.0) < 1e-12
    # no facets: empty result, like the reference
    empty = fem.FacetBasis(m, fem.ElementTetP1(), facets=np.zeros(0, dtype=np.int32))
    assert fem.BilinearForm(F["bmass"]).assemble(empty).nnz == 0
    with pytest.raises(ValueError, match="Incompatible"):
        fem.FacetBasis(m, fem.ElementTriP1())
    with pytest.raises(NotImplementedError):
        fem.FacetBasis(fem.MeshHex.init_tensor(x[:3], x[:3], x[:3]), fem.ElementHex1())
