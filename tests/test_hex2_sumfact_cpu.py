"""CPU check of the formulation behind the sum-factorised Hex2 kernel (csrc/skb_hex_sf.cu),
stage by stage in numpy with the very tables the host hands to the kernel
(skfem_b200/hex_sumfact.py):

    Jacobian columns from the trilinear map   ->   G_de = adj adj^T W / |det|  (6 components)
    T1 = sum_q1 G * pp[type1]   ->   T2 = sum_q2 T1 * pp[type2]   ->   B = sum_q3 T2 * pp[type3]
    A_ij = Bsym_ij + Boff_ij + Boff_ji

must reproduce the reference's assembled CSR (tests/golden/hex2_*.npz, real reference) to the
tolerance of the GPU test (rtol 1e-12)."""
import numpy as np
import pytest
from scipy.sparse import coo_matrix

import skfem_b200 as fem
from skfem_b200 import hex_sumfact
from cases import load
from product import mesh_from

COMBOS = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]


def _type(d, e, k):
    return (1 if d == k else 0) + (2 if e == k else 0)


def emulate(basis, tab, form):
    nq = tab["nq"]
    p, t = basis.mesh.p, basis.mesh.t
    nel = t.shape[1]
    x = p[:, t]                                              # (3, 8, nel)
    Xc = x[:, tab["vtx"].astype(np.int64).reshape(2, 2, 2), :]   # (3, a, b, c, nel)
    g = tab["g"]
    J0 = np.einsum('ibce,bq,cr->ieqr', Xc[:, 1] - Xc[:, 0], g, g)          # (q2, q3)
    J1 = np.einsum('iace,aq,cr->ieqr', Xc[:, :, 1] - Xc[:, :, 0], g, g)    # (q1, q3)
    J2 = np.einsum('iabe,aq,br->ieqr', Xc[:, :, :, 1] - Xc[:, :, :, 0], g, g)  # (q1, q2)
    J = np.empty((3, 3, nel, nq, nq, nq))
    J[:, 0] = J0[:, :, None, :, :]
    J[:, 1] = J1[:, :, :, None, :]
    J[:, 2] = J2[:, :, :, :, None]
    Jm = np.moveaxis(J, (0, 1), (-2, -1))                    # (nel, q1, q2, q3, 3, 3)
    det = np.linalg.det(Jm)
    adj = np.linalg.inv(Jm) * det[..., None, None]
    s = tab["qstride"].astype(np.int64)
    q1, q2, q3 = np.meshgrid(np.arange(nq), np.arange(nq), np.arange(nq), indexing='ij')
    W3 = basis.W[q1 * s[0] + q2 * s[1] + q3 * s[2]]
    pp = tab["pp"]
    if form == "mass":
        B = np.einsum('eqrs,aq,br,cs->eabc', np.abs(det) * W3, pp[0], pp[0], pp[0])
        Bs, Bo = B, None
    else:
        Bs = np.zeros((nel, 9, 9, 9))
        Bo = np.zeros((nel, 9, 9, 9))
        for d, e in COMBOS:
            G = np.einsum('eqrsj,eqrsj->eqrs', adj[..., d, :], adj[..., e, :]) * W3 / np.abs(det)
            B = np.einsum('eqrs,aq,br,cs->eabc', G, pp[_type(d, e, 0)], pp[_type(d, e, 1)],
                          pp[_type(d, e, 2)])
            if d == e:
                Bs += B
            else:
                Bo += B
    bn = tab["bnode"][:27].astype(np.int64)
    n1, n2, n3 = bn % 3, (bn // 3) % 3, bn // 9
    ia = 3 * n1[:, None] + n1[None, :]
    ib = 3 * n2[:, None] + n2[None, :]
    ic = 3 * n3[:, None] + n3[None, :]
    A = Bs[:, ia, ib, ic]
    if Bo is not None:
        A = A + (Bo[:, ia, ib, ic] + Bo[:, ia.T, ib.T, ic.T])
    return A                                                 # (nel, 27, 27)


@pytest.mark.parametrize("name", ["hex2_tensor2", "hex2_morphed2", "hex2_morphed4",
                                  "hex2_boxes3"])
def test_hex2_sum_factorisation_matches_reference(name):
    g = load(name)
    b = fem.Basis(mesh_from(g, "hex"), fem.ElementHex2())
    tab = hex_sumfact.tables(b)
    assert tab is not None and tab["nq"] == 7 and sorted(tab["qstride"]) == [1, 7, 49]
    edofs, N = g["element_dofs"], int(g["N"])
    nel = edofs.shape[1]
    for form in ("laplace", "mass"):
        local = emulate(b, tab, form)
        if form + "_local" in g.files:
            loc = np.moveaxis(g[form + "_local"].reshape(27, 27, nel), -1, 0)
            np.testing.assert_allclose(local, loc, rtol=1e-12, atol=1e-12 * np.abs(loc).max())
        rows = np.broadcast_to(edofs.T[:, :, None], (nel, 27, 27)).reshape(-1)
        cols = np.broadcast_to(edofs.T[:, None, :], (nel, 27, 27)).reshape(-1)
        A = coo_matrix((local.reshape(-1), (rows, cols)), shape=(N, N))
        A.eliminate_zeros()
        A = A.tocsr()
        assert np.array_equal(A.indptr, g[form + "_indptr"])
        assert np.array_equal(A.indices, g[form + "_indices"])
        ref = g[form + "_data"]
        np.testing.assert_allclose(A.data, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


def test_tables_are_refused_off_the_default_rule():
    x = np.linspace(0, 1, 3)
    m = fem.MeshHex.init_tensor(x, x, x)
    assert hex_sumfact.tables(fem.Basis(m, fem.ElementHex2(), intorder=4)) is None
    assert hex_sumfact.tables(fem.Basis(m, fem.ElementHex1())) is None
