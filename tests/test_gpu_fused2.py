"""GPU parity tests of the second-generation fused P1 path (csrc/skb_p1_fused2.cu,
skfem_b200/fused2.py) through the public API: warm re-assembly against the golden vectors of
the real reference and against the oracle on seeded unstructured inputs; pattern validation
when the mesh moves (basis.update_points).

Bar: indptr / indices bit-exact, CSR values within rtol 1e-12 (+ atol 1e-12 max|A|, SURVEY
A.9), repeated runs bit-identical."""
import numpy as np
import pytest

import skfem_b200 as fem
from cases import load, mesh_of
from product import mesh_from

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _oracle_csr(p, t, elements=None):
    from oracle import skfem_oracle as O
    return O.assemble_bilinear(O.laplace, O.cell_basis(mesh_of(dict(p=p, t=t), "tet"),
                                                       O.element("tet_p1"), elements=elements))


def _same(A, Ao):
    assert np.array_equal(A.indptr, Ao.indptr) and np.array_equal(A.indices, Ao.indices)
    np.testing.assert_allclose(A.data, Ao.data, rtol=RTOL, atol=RTOL * np.abs(Ao.data).max())


def _fused_plan(b):
    from skfem_b200.models.poisson import laplace
    fp = b._plans[("fused", laplace._plan_key(b, None, {}))]
    assert fp is not None and fp.version == 2
    return fp


@pytest.fixture
def options():
    from skfem_b200 import form as F
    saved = dict(F._CONFIG)
    yield F.set_options
    F._CONFIG.update(saved)


@pytest.mark.parametrize("tile,ring,pool,S", [(256, 3, 2048, None), (128, 3, 1024, None),
                                              (512, 3, 4096, None), (256, 4, 4096, 1),
                                              (128, 3, 2048, 2)])
def test_fused2_golden_and_oracle(options, tile, ring, pool, S):
    from skfem_b200.models.poisson import laplace
    options(fused=True, fused_version=2, fused2_tile=tile, fused2_ring=ring, fused2_pool=pool,
            fused2_S=S)
    for name in ["tet_p1_tensor6", "tet_p1_ball2", "tet_p1_refined3", "tet_p1_morphed5",
                 "tet_p1_tensor_nonuniform"]:
        g = load(name)
        b = fem.Basis(mesh_from(g, "tet"), fem.ElementTetP1())
        A0 = laplace.assemble(b)                      # cold, generic path
        A1 = laplace.assemble(b)                      # warm, fused path
        fp = _fused_plan(b)
        A2 = laplace.assemble(b)
        assert np.array_equal(A1.indptr, g["laplace_indptr"])
        assert np.array_equal(A1.indices, g["laplace_indices"])
        ref = g["laplace_data"]
        np.testing.assert_allclose(A1.data, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())
        assert np.array_equal(A1.data, A2.data)
        np.testing.assert_allclose(A1.data, A0.data, rtol=RTOL, atol=RTOL * np.abs(A0.data).max())
        assert not fp.flag.item()
    # several super-tiles, many shared slots, unstructured geometry
    rng = np.random.default_rng(3)
    x = np.sort(rng.random(21)); y = np.sort(rng.random(19)); z = np.sort(rng.random(20))
    m = fem.MeshTet.init_tensor(x, y, z)
    m = fem.MeshTet(m.p + 0.002 * rng.standard_normal(m.p.shape), m.t)
    b = fem.Basis(m, fem.ElementTetP1())
    laplace.assemble(b)
    A = laplace.assemble(b)
    _same(A, _oracle_csr(m.p, m.t))
    fp = _fused_plan(b)
    assert fp.ntiles == -(-m.nelements // fp.T) and fp.nst > 1 and fp.T <= tile
    assert fp.mode == 2 and not fp.flag.item()
    # element subset
    sub = np.arange(0, m.nelements, 3)
    bs = fem.Basis(m, fem.ElementTetP1(), elements=sub)
    laplace.assemble(bs)
    _same(laplace.assemble(bs), _oracle_csr(m.p, m.t, elements=sub))


def test_fused2_arithmetic_modes(options):
    """Coordinates outside [2^-28, 2^28] select the three-operation quadrature sum (mode 1),
    outside [2^-60, 2^60] plain IEEE division (mode 0); a 5-point rule with equal weights the
    generic sum.  All against the oracle on the same inputs."""
    from skfem_b200.models.poisson import laplace
    options(fused=True, fused_version=2, fused2_tile=128, fused2_pool=2048)
    rng = np.random.default_rng(11)
    x = np.sort(rng.random(9)); y = np.sort(rng.random(8)); z = np.sort(rng.random(10))
    m0 = fem.MeshTet.init_tensor(x, y, z)
    p0 = m0.p + 0.004 * rng.standard_normal(m0.p.shape)
    for scale, mode in ((1.0, 2), (2.0 ** 40, 1), (2.0 ** -80, 0)):
        m = fem.MeshTet(p0 * scale, m0.t)
        b = fem.Basis(m, fem.ElementTetP1())
        laplace.assemble(b)
        A = laplace.assemble(b)
        assert _fused_plan(b).mode == mode
        _same(A, _oracle_csr(m.p, m.t))
    # Kuhn grid: value-dependent 7-point pattern in all modes
    g = np.linspace(0, 1, 9)
    for scale, mode in ((1.0, 2), (2.0 ** 40, 1)):
        m = fem.MeshTet.init_tensor(g * scale, g * scale, g * scale)
        b = fem.Basis(m, fem.ElementTetP1())
        laplace.assemble(b)
        A = laplace.assemble(b)
        assert _fused_plan(b).mode == mode and A.nnz == 9 ** 3 + 6 * 8 * 9 ** 2
        _same(A, _oracle_csr(m.p, m.t))


def test_fused2_fast_arithmetic_within_tolerance(options):
    from skfem_b200.models.poisson import laplace
    options(fused=True, fused_version=2, fused2_tile=256)
    g = load("tet_p1_morphed5")
    b = fem.Basis(mesh_from(g, "tet"), fem.ElementTetP1())
    laplace.assemble(b)
    options(fused_arith="fast")
    A1, A2 = laplace.assemble(b), laplace.assemble(b)
    ref = g["laplace_data"]
    assert np.array_equal(A1.indices, g["laplace_indices"]) and np.array_equal(A1.data, A2.data)
    np.testing.assert_allclose(A1.data, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())


def test_update_points_revalidates_the_pattern(options):
    """Warm re-assembly after the mesh moved (basis.update_points): a motion that keeps every
    zero of the local matrices (anisotropic scaling of a Kuhn grid) reuses the plan and gives
    the oracle's matrix of the moved mesh; a motion that destroys them (shear + bending) is
    detected by the kernel's zero-mask check and re-planned - the result is again the
    oracle's CSR, now with the 15-point pattern; moving back restores the 7-point pattern."""
    from skfem_b200.models.poisson import laplace
    options(fused=True, fused_version=2, fused2_tile=128, fused2_pool=2048)
    g = np.linspace(0, 1, 11)
    m = fem.MeshTet.init_tensor(g, g, g)
    p0, t = m.p.copy(), m.t.copy()
    b = fem.Basis(m, fem.ElementTetP1())
    laplace.assemble(b)
    A = laplace.assemble(b)
    nnz7 = A.nnz
    fp = _fused_plan(b)
    # (1) anisotropic scaling: same zeros, new values
    p1 = p0 * np.array([[1.3], [0.7], [2.1]])
    b.update_points(p1)
    A1 = laplace.assemble(b)
    assert _fused_plan(b) is fp                        # plan reused
    assert A1.nnz == nnz7
    _same(A1, _oracle_csr(p1, t))
    assert np.abs(A1.data - A.data).max() > 1e-3       # values really changed
    # (2) shear + bending: exact zeros disappear -> detected, re-planned
    p2 = p0.copy()
    p2[0] = p0[0] + 0.03 * np.sin(7 * p0[1])
    p2[1] = p0[1] + 0.02 * p0[2] ** 2
    b.update_points(p2)
    A2 = laplace.assemble(b)
    assert A2.nnz > nnz7
    _same(A2, _oracle_csr(p2, t))
    A2w = laplace.assemble(b)                          # warm again on the new plan
    _same(A2w, _oracle_csr(p2, t))
    assert _fused_plan(b) is not fp
    # (3) back to the grid: nonzeros become exact zeros -> detected again
    b.update_points(p0)
    A3 = laplace.assemble(b)
    assert A3.nnz == nnz7
    _same(A3, _oracle_csr(p0, t))
    # (4) with out= (graph-style loop) a changed pattern cannot be returned in place
    import torch
    laplace.assemble(b)
    out = torch.empty(nnz7, dtype=torch.float64, device="cuda")
    laplace.assemble_device(b, out=out)
    b.update_points(p2)
    from skfem_b200.form import PatternChanged
    with pytest.raises(PatternChanged):
        laplace.assemble_device(b, out=out)


def test_fused2_mass_form(options):
    """The fused path also takes the mass form u * v on ElementTetP1 (4-point rule): golden
    vectors of the real reference, the oracle on an unstructured mesh with several
    super-tiles, bit-identical repeats, re-assembly on moved points."""
    from oracle import skfem_oracle as O
    from skfem_b200.models.poisson import mass
    options(fused=True, fused_version=2, fused2_tile=128, fused2_pool=2048)

    def plan_of(b):
        fp = b._plans[("fused", mass._plan_key(b, None, {}))]
        assert fp is not None and fp.version == 2 and fp.mode == 4
        return fp
    for name in ["tet_p1_tensor6", "tet_p1_ball2", "tet_p1_morphed5", "tet_p1_tensor_nonuniform"]:
        g = load(name)
        b = fem.Basis(mesh_from(g, "tet"), fem.ElementTetP1())
        A0 = mass.assemble(b)                         # cold, generic path
        A1 = mass.assemble(b)                         # warm, fused path
        plan_of(b)
        A2 = mass.assemble(b)
        assert np.array_equal(A1.indptr, g["mass_indptr"])
        assert np.array_equal(A1.indices, g["mass_indices"])
        ref = g["mass_data"]
        np.testing.assert_allclose(A1.data, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())
        assert np.array_equal(A1.data, A2.data)
        np.testing.assert_allclose(A1.data, A0.data, rtol=RTOL, atol=RTOL * np.abs(A0.data).max())
    rng = np.random.default_rng(5)
    x = np.sort(rng.random(17)); y = np.sort(rng.random(15)); z = np.sort(rng.random(16))
    m = fem.MeshTet.init_tensor(x, y, z)
    p0 = m.p + 0.002 * rng.standard_normal(m.p.shape)
    m = fem.MeshTet(p0, m.t)
    b = fem.Basis(m, fem.ElementTetP1())
    mass.assemble(b)
    A = mass.assemble(b)
    fp = plan_of(b)
    assert fp.nst > 1

    def oracle(p):
        return O.assemble_bilinear(O.mass, O.cell_basis(mesh_of(dict(p=p, t=m.t), "tet"),
                                                        O.element("tet_p1")))
    _same(A, oracle(p0))
    assert abs(A.sum() - np.abs(np.linalg.det(
        (p0[:, m.t[1:]] - p0[:, m.t[:1]]).transpose(2, 0, 1))).sum() / 6.0) < 1e-12
    p1 = p0 * np.array([[1.5], [0.6], [1.1]])
    b.update_points(p1)
    A1 = mass.assemble(b)
    assert plan_of(b) is fp and not fp.flag.item()
    _same(A1, oracle(p1))


def test_fused2_baseline_config2_full_size(options):
    """BASELINE configs[1] at full size through the v2 path: closed-form nnz, symmetry, zero
    row sums, exact energy of a linear field, agreement of the cold and warm paths, and a
    moved mesh (anisotropic scaling) against the scaled closed form."""
    from skfem_b200.models.poisson import laplace
    options(fused=True, fused_version=2)
    n = 100
    x = np.linspace(0, 1, n + 1)
    b = fem.Basis(fem.MeshTet.init_tensor(x, x, x), fem.ElementTetP1())
    A0 = laplace.assemble(b)
    A1 = laplace.assemble(b)
    assert _fused_plan(b).mode == 2
    assert A0.nnz == 7150901
    assert np.array_equal(A0.indptr, A1.indptr) and np.array_equal(A0.indices, A1.indices)
    scale = np.abs(A0.data).max()
    np.testing.assert_allclose(A1.data, A0.data, rtol=1e-12, atol=1e-12 * scale)
    assert abs(A1 - A1.T).max() == 0.0
    assert np.abs(A1 @ np.ones(b.N)).max() <= 1e-11 * scale
    u = 2.0 * b.mesh.p[0] - 3.0 * b.mesh.p[1] + 0.5 * b.mesh.p[2]   # |grad u|^2 = 13.25
    np.testing.assert_allclose(u @ (A1 @ u), 13.25, rtol=1e-11)
    s = np.array([[1.25], [0.8], [1.1]])
    b.update_points(b.mesh.p * s)
    A2 = laplace.assemble(b)
    assert np.array_equal(A2.indices, A0.indices)
    u = 2.0 * b.mesh.p[0] - 3.0 * b.mesh.p[1] + 0.5 * b.mesh.p[2]
    np.testing.assert_allclose(u @ (A2 @ u), 13.25 * float(np.prod(s)), rtol=1e-11)


def test_full_size_c2_against_the_real_reference(options):
    """BASELINE configs[1] at full size (6.0 M tets) against the UNMODIFIED reference shipped
    in oracle/_ref (tools/install_ref.sh): indptr / indices bit-exact, values rtol 1e-12, for
    the cold (generic) and the warm (fused) path.  About 30 s and 8 GB of host memory."""
    import os
    import sys
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref, "skfem")):
        pytest.skip("oracle/_ref not installed (tools/install_ref.sh)")
    sys.path.insert(0, ref)
    try:
        import skfem
        from skfem.models.poisson import laplace as ref_laplace
        n = 100
        x = np.linspace(0, 1, n + 1)
        mr = skfem.MeshTet.init_tensor(x, x, x)
        Ar = skfem.BilinearForm(ref_laplace.form, nthreads=min(os.cpu_count() or 1, 8)).assemble(
            skfem.Basis(mr, skfem.ElementTetP1()))
    finally:
        sys.path.remove(ref)
    from skfem_b200.models.poisson import laplace
    options(fused=True, fused_version=2)
    m = fem.MeshTet.init_tensor(x, x, x)
    assert np.array_equal(m.p, mr.p) and np.array_equal(m.t, mr.t)
    b = fem.Basis(m, fem.ElementTetP1())
    for A in (laplace.assemble(b), laplace.assemble(b)):      # cold, then warm (fused)
        assert A.nnz == Ar.nnz == 7150901
        assert np.array_equal(A.indptr, Ar.indptr) and np.array_equal(A.indices, Ar.indices)
        np.testing.assert_allclose(A.data, Ar.data, rtol=1e-12, atol=1e-12 * np.abs(Ar.data).max())
    assert _fused_plan(b) is not None
