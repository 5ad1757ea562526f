"""Device-side boundary conditions: ``enforce`` / ``condense``
/ ``solve`` on the device CSR.  Golden vectors tests/golden/bc_*.npz come from
the real reference (tools/gen_golden_bc.py: skfem.utils.enforce / condense /
solve).  CPU: the oracle's entry-by-entry restatement is pinned bitwise.  GPU:
the product's kernels (csrc/skb_bc.cu) reproduce the same arrays bitwise from a
matrix assembled on the device."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix

from cases import load

CASES = [("bc_tri_p1", "MeshTri", "ElementTriP1"), ("bc_tet_p2", "MeshTet", "ElementTetP2")]


def gcsr(g, prefix):
    return csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]),
                      shape=tuple(g[prefix + "_shape"]))


def same_csr(A, g, prefix):
    A = A.tocsr()
    assert A.shape == tuple(g[prefix + "_shape"])
    assert np.array_equal(A.indptr, g[prefix + "_indptr"])
    assert np.array_equal(A.indices, g[prefix + "_indices"])
    assert np.array_equal(A.data, g[prefix + "_data"])


@pytest.mark.parametrize("name,M,E", CASES)
def test_oracle_bc(name, M, E):
    from oracle import skfem_oracle as O
    g = load(name)
    A, D, x, b = gcsr(g, "A"), g["D"], g["x"], g["b"]
    Ae, be = O.enforce(A, b, D=D)
    same_csr(Ae, g, "enf")
    assert np.array_equal(be, g["enf_b"])
    Ae, be = O.enforce(A, b, x=x, D=D, diag=2.5)
    same_csr(Ae, g, "enf2")
    assert np.array_equal(be, g["enf2_b"])
    AII, bI, I = O.condense(A, b, D=D)
    same_csr(AII, g, "con")
    assert np.array_equal(bI, g["con_b"]) and np.array_equal(I, g["con_I"])
    AII, bI, I = O.condense(A, b, x=x, D=D)
    same_csr(AII, g, "con2")
    assert np.array_equal(bI, g["con2_b"])
    assert np.array_equal(O.condense(A, None, x=x, D=D)[1], g["con5_b"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,M,E", CASES)
def test_gpu_bc(name, M, E):
    import torch
    import skfem_b200 as fem
    from skfem_b200.models.poisson import laplace, mass, unit_load
    g = load(name)
    m = getattr(fem, M)(g["p"], g["t"])
    basis = fem.Basis(m, getattr(fem, E)())
    D, x = g["D"], g["x"]
    assert np.array_equal(np.sort(basis.get_dofs()), np.sort(D))
    # bitwise checks start from the reference's own matrix (CSR values of an
    # assembled matrix agree only to rtol 1e-12) ...
    A = fem.DeviceCSR.from_scipy(gcsr(g, "A"))
    b = torch.as_tensor(g["b"], device=A.data.device)
    xd = torch.as_tensor(x, device=A.data.device)
    Ae, be = fem.enforce(A, b, D=D)
    same_csr(Ae.to_scipy(), g, "enf")
    assert np.array_equal(be.cpu().numpy(), g["enf_b"])
    assert np.array_equal(A.data.cpu().numpy(), g["A_data"])          # not overwritten
    Ae, be = fem.enforce(A, b, x=xd, D=D, diag=2.5)
    same_csr(Ae.to_scipy(), g, "enf2")
    assert np.array_equal(be.cpu().numpy(), g["enf2_b"])
    AII, bI, xI, I = fem.condense(A, b, D=D)
    same_csr(AII.to_scipy(), g, "con")
    assert np.array_equal(bI.cpu().numpy(), g["con_b"])
    assert np.array_equal(I.cpu().numpy(), g["con_I"])
    AII, bI, xI, I = fem.condense(A, b, x=xd, D=D)
    same_csr(AII.to_scipy(), g, "con2")
    assert np.array_equal(bI.cpu().numpy(), g["con2_b"])
    AII3, bI3 = fem.condense(A, b, x=xd, I=g["con_I"], expand=False)
    same_csr(AII3.to_scipy(), g, "con2")
    assert np.array_equal(bI3.cpu().numpy(), g["con2_b"])
    assert np.array_equal(fem.condense(A, x=xd, D=D, expand=False)[1].cpu().numpy(), g["con5_b"])
    # ... matrix right-hand sides (eigenvalue problems) from the device-assembled mass matrix
    Md = mass.assemble_device(basis)
    Mref = Md.to_scipy()
    _, Me = fem.enforce(A, Md, D=D)
    Mexp = Mref.copy()
    rows = np.repeat(np.arange(Mref.shape[0]), np.diff(Mref.indptr))
    Mexp.data[np.isin(rows, D)] = 0.
    assert np.array_equal(Me.to_scipy().data, Mexp.data)
    _, MII, _, _ = fem.condense(A, Md, D=D)
    I_ = g["con_I"]
    ref = Mref[I_][:, I_]
    got = MII.to_scipy()
    assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
    assert np.array_equal(got.data, ref.data)
    # ... and the whole pipeline on the device: assemble -> condense -> solve
    Ad = laplace.assemble_device(basis)
    bd = unit_load.assemble_device(basis)
    sol = fem.solve(*fem.condense(Ad, bd, D=D)).cpu().numpy()
    np.testing.assert_allclose(sol, g["sol"], rtol=0, atol=1e-9 * np.abs(g["sol"]).max())
    sol2 = fem.solve(*fem.condense(Ad, bd, x=xd, D=D)).cpu().numpy()
    np.testing.assert_allclose(sol2, g["sol2"], rtol=0, atol=1e-9 * np.abs(g["sol2"]).max())
    # spmv in scipy's order
    y = fem.utils.matvec(A, xd).cpu().numpy()
    assert np.array_equal(y, gcsr(g, "A") @ x)
    with pytest.raises(Exception, match="Either I or D"):
        fem.condense(A, b)
    with pytest.raises(TypeError):
        fem.condense(gcsr(g, "A"), g["b"], D=D)


@pytest.mark.gpu
def test_gpu_ex01_on_device():
    """docs/examples/ex01.py end to end on the device against the golden solution."""
    import skfem_b200 as fem
    from skfem_b200.models.poisson import laplace, unit_load
    g = load("ex01_solution")
    m = fem.MeshTri().refined(int(g["refines"]) if "refines" in g.files else 4)
    basis = fem.Basis(m, fem.ElementTriP1())
    A = laplace.assemble_device(basis)
    b = unit_load.assemble_device(basis)
    x = fem.solve(*fem.condense(A, b, D=basis.get_dofs())).cpu().numpy()
    key = "x" if "x" in g.files else [k for k in g.files if g[k].shape == x.shape][0]
    np.testing.assert_allclose(x, g[key], rtol=0, atol=1e-10)


@pytest.mark.gpu
def test_gpu_condense_with_named_sets_sharing_dofs():
    """D given as a dict of named sets that share corner DOFs is flattened with
    np.unique(np.concatenate(...)) like the reference (utils.py:277-288)."""
    import skfem_b200 as fem
    from skfem_b200.models.poisson import laplace, unit_load
    g = load("bc_tri_p1")
    m = fem.MeshTri(g["p"], g["t"])
    basis = fem.Basis(m, fem.ElementTriP1())
    A, b = laplace.assemble_device(basis), unit_load.assemble_device(basis)
    left = np.nonzero(m.p[0] == m.p[0].min())[0]
    bottom = np.nonzero(m.p[1] == m.p[1].min())[0]
    assert np.intersect1d(left, bottom).size > 0
    both = np.unique(np.concatenate([left, bottom]))
    A1, b1, x1, I1 = fem.condense(A, b, D={"left": left, "bottom": bottom})
    A2, b2, x2, I2 = fem.condense(A, b, D=both)
    S1, S2 = A1.to_scipy(), A2.to_scipy()
    assert np.array_equal(S1.indptr, S2.indptr) and np.array_equal(S1.indices, S2.indices)
    assert np.array_equal(S1.data, S2.data)
    assert np.array_equal(b1.cpu().numpy(), b2.cpu().numpy())
    assert np.array_equal(np.asarray(I1.cpu() if hasattr(I1, "cpu") else I1),
                          np.asarray(I2.cpu() if hasattr(I2, "cpu") else I2))
