"""Pin the oracle (oracle/skfem_oracle.py) against vectors produced by the
real reference (tools/gen_golden.py).  Bitwise for local data, CSR structure
and load vectors."""
import numpy as np
import pytest

from oracle import skfem_oracle as O
from cases import CASES, LAME, load, mesh_of


def _forms():
    def user_aniso(u, v, w):
        return (1. + w.x[0] * w.x[1]) * O.dot(O.grad(u), O.grad(v)) + 3. * u * v

    def user_load(v, w):
        return np.sin(3. * w.x[0]) * v + w.x[1] * v

    return dict(laplace=O.laplace, mass=O.mass, vector_laplace=O.vector_laplace,
                elasticity=O.linear_elasticity(*LAME), user_aniso=user_aniso,
                unit_load=O.unit_load, user_load=user_load)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference(name):
    refdom, ename, vector, bil, lin, has_local = CASES[name]
    g = load(name)
    m = mesh_of(g, refdom)
    e = O.element(ename, vector=vector)
    b = O.cell_basis(m, e)
    assert np.array_equal(b.element_dofs, g["element_dofs"])
    assert b.element_dofs.dtype == np.int32
    assert b.N == int(g["N"])
    assert np.array_equal(b.X, g["X"]) and np.array_equal(b.W, g["W"])
    forms = _forms()
    for f in bil:
        form = forms[f]
        if f == "mass" and vector:
            form = lambda u, v, w: O.dot(u, v)  # noqa: E731
        idx, data, shape = O.bilinear_coo(form, b)
        assert np.array_equal(data, g[f + "_local"]), f
        A = O.coo_to_csr(idx, data, shape)
        assert np.array_equal(A.indptr, g[f + "_indptr"])
        assert np.array_equal(A.indices, g[f + "_indices"])
        assert np.array_equal(A.data, g[f + "_data"])
    for f in lin:
        vec = O.assemble_linear(forms[f], b)
        assert np.array_equal(vec, g[f + "_vec"]), f


def test_oracle_chunked_equals_monolithic():
    g = load("tet_p1_morphed5")
    m = mesh_of(g, "tet")
    A = O.assemble_bilinear_chunked(O.laplace, m, O.element("tet_p1"), chunk=97)
    assert np.array_equal(A.indptr, g["laplace_indptr"])
    assert np.array_equal(A.indices, g["laplace_indices"])
    np.testing.assert_allclose(A.data, g["laplace_data"], rtol=1e-12, atol=0)


def test_oracle_mesh_builders_and_known_answers():
    # C1: MeshTri().refined(4) rebuilt by the oracle's own constructors
    g = load("c1_tri_p1_refined4")
    m = O.refine_tri(O.mesh_tri_default(), 4)
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])
    g = load("tet_p1_tensor6")
    lin = np.linspace(0, 1, 7)
    m = O.mesh_tet_tensor(lin, lin, lin)
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])
    g = load("hex1_tensor3")
    lin = np.linspace(0, 1, 4)
    m = O.mesh_hex_tensor(lin, lin, lin)
    assert np.array_equal(m.p, g["p"]) and np.array_equal(m.t, g["t"])
    # doctest known answer, skfem/assembly/__init__.py:38-46
    b = O.cell_basis(O.mesh_tri_default(), O.element("tri_p1"))
    M = O.assemble_bilinear(O.mass, b).toarray()
    ref = np.array([[0.08333333, 0.04166667, 0.04166667, 0.],
                    [0.04166667, 0.16666667, 0.08333333, 0.04166667],
                    [0.04166667, 0.08333333, 0.16666667, 0.04166667],
                    [0., 0.04166667, 0.04166667, 0.08333333]])
    np.testing.assert_allclose(M, ref, atol=1e-8)
    f = O.assemble_linear(O.unit_load, b)
    np.testing.assert_allclose(f, [1 / 6, 1 / 3, 1 / 3, 1 / 6], atol=1e-12)
    # closed-form nnz of the P1 Laplace 7-point stencil
    n = 6
    b = O.cell_basis(O.mesh_tet_tensor(*(3 * (np.linspace(0, 1, n + 1),))),
                     O.element("tet_p1"))
    A = O.assemble_bilinear(O.laplace, b)
    assert A.nnz == (n + 1) ** 3 + 6 * n * (n + 1) ** 2


def test_oracle_pairwise_sum_is_numpy_order():
    rng = np.random.default_rng(0)
    for n in [1, 3, 4, 7, 8, 9, 11, 15, 16, 17, 64, 100, 128, 129, 200, 343]:
        a = rng.standard_normal((5, n)) * 10.0 ** rng.integers(-8, 8, (5, n))
        ref = np.sum(a, axis=1)
        got = np.array([O.pairwise_sum(r) for r in a])
        assert np.array_equal(ref, got), n
