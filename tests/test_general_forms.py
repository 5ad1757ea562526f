"""skfem.models.general.divu with mixed (vector-P2 trial, P1 test) bases and
skfem.helpers.inv / det / mul / identity inside a user form with a vector
coefficient field.  Golden vectors: tests/golden/general_forms.npz (real
reference, tools/gen_golden_general.py)."""
import numpy as np
import pytest

from cases import load, mesh_of


def _deformed(H):
    def deformed_laplace(u, v, w):
        F = H.grad(w['disp']) + H.identity(w['disp'])
        Finv = H.inv(F)
        return H.dot(H.mul(Finv, H.grad(u)), H.mul(Finv, H.grad(v))) * H.det(F)
    return deformed_laplace


def test_oracle_general_forms():
    from oracle import skfem_oracle as O
    g = load("general_forms")
    m = mesh_of(g, "tet")
    ub = O.cell_basis(m, O.element("tet_p2", vector=True))
    pb = O.cell_basis(m, O.element("tet_p1"), intorder=4)
    idx, data, shape = O.bilinear_coo(O.divu, ub, vbasis=pb)
    assert shape == tuple(g["divu_shape"])
    assert np.array_equal(data, g["divu_local"])
    A = O.coo_to_csr(idx, data, shape)
    assert np.array_equal(A.indptr, g["divu_indptr"]) and np.array_equal(A.indices, g["divu_indices"])
    assert np.array_equal(A.data, g["divu_data"])
    idx, data, shape = O.bilinear_coo(_deformed(O), pb, disp=O.interpolate(ub, g["disp"]))
    assert np.array_equal(data, g["defo_local"])
    A = O.coo_to_csr(idx, data, shape)
    assert np.array_equal(A.indices, g["defo_indices"]) and np.array_equal(A.data, g["defo_data"])


@pytest.mark.gpu
def test_gpu_general_forms():
    import skfem_b200 as fem
    from skfem_b200 import helpers as H
    from skfem_b200.models.general import divu
    g = load("general_forms")
    m = fem.MeshTet(g["p"], g["t"])
    ub = fem.Basis(m, fem.ElementVector(fem.ElementTetP2()))
    pb = fem.Basis(m, fem.ElementTetP1(), intorder=4)
    assert np.array_equal(divu.elemental(ub, pb).data, g["divu_local"])
    B = divu.assemble(ub, pb)
    assert B.shape == tuple(g["divu_shape"])
    assert np.array_equal(B.indptr, g["divu_indptr"]) and np.array_equal(B.indices, g["divu_indices"])
    np.testing.assert_allclose(B.data, g["divu_data"], rtol=1e-12,
                               atol=1e-12 * np.abs(g["divu_data"]).max())
    form = fem.BilinearForm(_deformed(H))
    disp = ub.interpolate(g["disp"])
    loc = form.elemental(pb, disp=disp).data
    np.testing.assert_allclose(loc, g["defo_local"], rtol=1e-13,
                               atol=1e-15 * np.abs(g["defo_local"]).max())
    A = form.assemble(pb, disp=disp)
    assert np.array_equal(A.indptr, g["defo_indptr"]) and np.array_equal(A.indices, g["defo_indices"])
    np.testing.assert_allclose(A.data, g["defo_data"], rtol=1e-12,
                               atol=1e-12 * np.abs(g["defo_data"]).max())


def _cross_form(H):
    def cross_form(u, v, w):
        return H.dot(H.cross(u, w['w']), v)
    return cross_form


def test_oracle_curl_forms():
    """models.general.curluv / rot / vrot and helpers.cross (these build their
    result with np.array([...]), which numpy does not dispatch for device fields)."""
    from oracle import skfem_oracle as O
    g = load("general_curl")
    m = mesh_of(g, "tet")
    vb = O.cell_basis(m, O.element("tet_p1", vector=True))
    wf = O.interpolate(vb, g["wdofs"])
    idx, data, shape = O.bilinear_coo(O.curluv, vb)
    assert np.array_equal(data, g["curluv_local"])
    A = O.coo_to_csr(idx, data, shape)
    assert np.array_equal(A.indptr, g["curluv_indptr"]) and np.array_equal(A.indices, g["curluv_indices"])
    assert np.array_equal(O.assemble_linear(O.rot, vb, w=wf), g["rot_vec"])
    assert np.array_equal(O.assemble_linear(O.vrot, vb, w=wf), g["vrot_vec"])
    idx, data, shape = O.bilinear_coo(_cross_form(O), vb, w=wf)
    assert np.array_equal(data, g["cross_local"])
    m2 = mesh_of(dict(p=g["p2"], t=g["t2"]), "tri")
    vb2 = O.cell_basis(m2, O.element("tri_p1", vector=True))
    sb2 = O.cell_basis(m2, O.element("tri_p1"))
    idx, data, shape = O.bilinear_coo(O.curluv, vb2, vbasis=sb2)
    assert shape == tuple(g["curluv2_shape"]) and np.array_equal(data, g["curluv2_local"])


@pytest.mark.gpu
def test_gpu_curl_forms():
    import skfem_b200 as fem
    from skfem_b200 import helpers as H
    from skfem_b200.models.general import curluv, rot, vrot
    g = load("general_curl")
    m = fem.MeshTet(g["p"], g["t"])
    vb = fem.Basis(m, fem.ElementVector(fem.ElementTetP1()))
    wf = vb.interpolate(g["wdofs"])
    assert np.array_equal(curluv.elemental(vb).data, g["curluv_local"])
    A = curluv.assemble(vb)
    assert np.array_equal(A.indptr, g["curluv_indptr"]) and np.array_equal(A.indices, g["curluv_indices"])
    np.testing.assert_allclose(A.data, g["curluv_data"], rtol=1e-12,
                               atol=1e-12 * np.abs(g["curluv_data"]).max())
    assert np.array_equal(rot.assemble(vb, w=wf), g["rot_vec"])
    assert np.array_equal(vrot.assemble(vb, w=wf), g["vrot_vec"])
    form = fem.BilinearForm(_cross_form(H))
    assert np.array_equal(form.elemental(vb, w=wf).data, g["cross_local"])
    C = form.assemble(vb, w=wf)
    assert np.array_equal(C.indptr, g["cross_indptr"]) and np.array_equal(C.indices, g["cross_indices"])
    np.testing.assert_allclose(C.data, g["cross_data"], rtol=1e-12,
                               atol=1e-12 * np.abs(g["cross_data"]).max())
    m2 = fem.MeshTri(g["p2"], g["t2"])
    vb2 = fem.Basis(m2, fem.ElementVector(fem.ElementTriP1()))
    sb2 = fem.Basis(m2, fem.ElementTriP1())
    assert np.array_equal(curluv.elemental(vb2, sb2).data, g["curluv2_local"])
    A2 = curluv.assemble(vb2, sb2)
    assert A2.shape == tuple(g["curluv2_shape"])
    assert np.array_equal(A2.indptr, g["curluv2_indptr"]) and np.array_equal(A2.indices, g["curluv2_indices"])
