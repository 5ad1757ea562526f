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
