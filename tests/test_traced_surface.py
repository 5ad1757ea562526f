"""The surface user-defined forms see: coefficient
fields from ``basis.interpolate`` passed as kwargs, ``Functional``, trial and
test functions from two different bases, ``asm`` with a bare callable.

Golden vectors: tests/golden/traced_surface.npz (real reference,
tools/gen_golden.py).  The CPU test pins the oracle; the GPU test checks the
product's traced path."""
import numpy as np
import pytest

from cases import load, mesh_of


def _oracle_forms(O):
    def newton_like(u, v, w):
        return (1. + w['prev'] ** 2) * O.dot(O.grad(u), O.grad(v)) \
            + O.dot(w['prev'].grad, O.grad(v)) * u

    def residual_like(v, w):
        return O.dot(w['prev'].grad, O.grad(v)) + w['prev'] * v * w['scale']

    def energy(w):
        return 0.5 * O.dot(w['prev'].grad, w['prev'].grad) + w.x[0] * w['prev']

    def mixed(u, v, w):
        return u * v + O.dot(O.grad(u), O.grad(v))

    def scaled_mass(u, v, w):
        return w['alpha'] * u * v
    return newton_like, residual_like, energy, mixed, scaled_mass


def test_oracle_traced_surface():
    from oracle import skfem_oracle as O
    g = load("traced_surface")
    m = mesh_of(g, "tet")
    newton_like, residual_like, energy, mixed, scaled_mass = _oracle_forms(O)
    b1 = O.cell_basis(m, O.element("tet_p1"))
    b2 = O.cell_basis(m, O.element("tet_p2"), intorder=2)
    prev = g["prev"]
    f = O.interpolate(b1, prev)
    assert np.array_equal(np.array(f), g["interp_value"])
    assert np.array_equal(f.grad, g["interp_grad"])
    idx, data, shape = O.bilinear_coo(newton_like, b1, prev=prev)
    assert np.array_equal(data, g["newton_local"])
    A = O.coo_to_csr(idx, data, shape)
    assert np.array_equal(A.indptr, g["newton_indptr"]) and np.array_equal(A.data, g["newton_data"])
    assert np.array_equal(O.assemble_linear(residual_like, b1, prev=prev, scale=2.5),
                          g["residual_vec"])
    el = O.functional_elemental(energy, b1, prev=prev)
    assert np.array_equal(el, g["energy_elemental"])
    assert np.sum(el) == float(g["energy"])
    idx, data, shape = O.bilinear_coo(mixed, b2, vbasis=b1)
    assert shape == tuple(g["mixed_shape"])
    assert np.array_equal(data, g["mixed_local"])
    A = O.coo_to_csr(idx, data, shape)
    assert np.array_equal(A.indices, g["mixed_indices"]) and np.array_equal(A.data, g["mixed_data"])
    A = O.assemble_bilinear(scaled_mass, b1, alpha=3.0)
    assert np.array_equal(A.data, g["asm_data"])


@pytest.mark.gpu
def test_gpu_traced_surface():
    import skfem_b200 as fem
    from skfem_b200.helpers import dot, grad
    g = load("traced_surface")
    m = fem.MeshTet(g["p"], g["t"])
    b1 = fem.Basis(m, fem.ElementTetP1())
    b2 = fem.Basis(m, fem.ElementTetP2(), intorder=2)
    prev = g["prev"]
    np.testing.assert_array_equal(b1.doflocs, g["doflocs"])

    @fem.BilinearForm
    def newton_like(u, v, w):
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

    f = b1.interpolate(prev)
    assert np.array_equal(f.numpy(), g["interp_value"])
    assert np.array_equal(f.grad.numpy(), g["interp_grad"])
    coo = newton_like.elemental(b1, prev=prev)
    assert np.array_equal(coo.data, g["newton_local"])
    A = newton_like.assemble(b1, prev=prev)
    assert np.array_equal(A.indptr, g["newton_indptr"])
    assert np.array_equal(A.indices, g["newton_indices"])
    np.testing.assert_allclose(A.data, g["newton_data"], rtol=1e-12,
                               atol=1e-12 * np.abs(g["newton_data"]).max())
    assert np.array_equal(residual_like.assemble(b1, prev=prev, scale=2.5), g["residual_vec"])
    el = energy.elemental(b1, prev=prev)
    assert np.array_equal(el, g["energy_elemental"])
    np.testing.assert_allclose(energy.assemble(b1, prev=prev), float(g["energy"]), rtol=1e-14)
    coo = mixed.elemental(b2, b1)
    assert np.array_equal(coo.data, g["mixed_local"])
    A = mixed.assemble(b2, b1)
    assert A.shape == tuple(g["mixed_shape"])
    assert np.array_equal(A.indptr, g["mixed_indptr"])
    assert np.array_equal(A.indices, g["mixed_indices"])
    np.testing.assert_allclose(A.data, g["mixed_data"], rtol=1e-12,
                               atol=1e-12 * np.abs(g["mixed_data"]).max())
    A = fem.asm(scaled_mass, b1, alpha=3.0)
    assert np.array_equal(A.indices, g["asm_indices"])
    np.testing.assert_allclose(A.data, g["asm_data"], rtol=1e-12)
    # asm hands the basis position to the form as w.idx, also for a single basis
    # (assembly/__init__.py:91-93)
    seen = []

    def idx_mass(u, v, w):
        seen.append(w['idx'])
        return 3.0 * u * v
    A2 = fem.asm(idx_mass, b1)
    assert seen and all(ix == (0,) for ix in seen)
    np.testing.assert_allclose(A2.data, g["asm_data"], rtol=1e-12)
    with pytest.raises(ValueError, match="wrong size"):
        b1.interpolate(prev[:-1])
