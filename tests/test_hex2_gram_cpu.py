"""CPU check of the formulation behind the Hex2 tensor-core kernel
(csrc/skb_hex_mma.cu): the element-local Laplace / mass matrix written as the Gram
matrix  A = Gs^T Gs  with  Gs[(q, d), i] = (invDF^T dphi_i)_d(q) * sqrt(dx_q)  (laplace)
or  Gs[q, i] = phi_i(q) * sqrt(dx_q)  (mass), zero-padded to 32 basis functions and to
whole chunks of 32 quadrature points, must reproduce the reference's assembled CSR
(tests/golden/hex2_*.npz, real reference) to the same tolerance the GPU test uses.
Geometry comes from the oracle's isoparametric restatement, the tables from the
product's host-side ElementHex2."""
import numpy as np
import pytest
from scipy.sparse import coo_matrix

import skfem_b200 as fem
from cases import load, mesh_of


@pytest.mark.parametrize("name", ["hex2_tensor2", "hex2_morphed2", "hex2_boxes3"])
def test_hex2_gram_formulation_matches_reference(name):
    from oracle import skfem_oracle as O
    g = load(name)
    m = mesh_of(g, "hex")
    X, W, edofs, N = g["X"], g["W"], g["element_dofs"], int(g["N"])
    phi, dphi = fem.ElementHex2().tabulate(X)            # (27, nqp), (27, 3, nqp)
    geo = O.iso_geometry(m, X)
    nel, nqp = geo.nel, W.shape[0]
    sdx = np.sqrt(np.abs(geo.detDF) * W[None, :])        # (nel, nqp)
    qpad = -(-nqp // 32) * 32                            # whole chunks of 32 points
    for form in ("laplace", "mass"):
        kd = 3 if form == "laplace" else 1
        Gs = np.zeros((nel, qpad * kd, 32))              # [element][k = q * kd + d][i]
        for i in range(27):
            if form == "laplace":
                # grad_j = sum_c invDF[c, j] dphi_i[c]     (element_h1.py:17)
                grad = np.einsum('cjeq,cq->jeq', geo.invDF, dphi[i])
                for d in range(3):
                    Gs[:, np.arange(nqp) * 3 + d, i] = grad[d] * sdx
            else:
                Gs[:, np.arange(nqp), i] = phi[i][None, :] * sdx
        local = np.einsum('eki,ekj->eij', Gs, Gs)[:, :27, :27]      # the DMMA contraction
        rows = np.broadcast_to(edofs.T[:, :, None], (nel, 27, 27)).reshape(-1)   # test dof i
        cols = np.broadcast_to(edofs.T[:, None, :], (nel, 27, 27)).reshape(-1)   # trial dof j
        A = coo_matrix((local.reshape(-1), (rows, cols)), shape=(N, N))
        A.eliminate_zeros()
        A = A.tocsr()
        assert np.array_equal(A.indptr, g[form + "_indptr"])
        assert np.array_equal(A.indices, g[form + "_indices"])
        ref = g[form + "_data"]
        np.testing.assert_allclose(A.data, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
