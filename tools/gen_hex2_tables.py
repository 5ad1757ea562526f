"""Reference tables of ElementHex2 at its default quadrature rule.

The reference evaluates ElementHex2.lbasis with machine-generated Horner forms
(skfem/element/element_hex/element_hex2.py:11-1210); re-deriving the triquadratic Lagrange
basis as a tensor product reproduces them only to a few ulp (<= 2e-14).  SURVEY Appendix A.3
asks for the reference's own numbers, so - exactly like the simplex quadrature tables
(tools/gen_quadrature_tables.py) - the values phi_b(X_q), dphi_b(X_q) at the default rule
(intorder 2 * maxdeg = 12: 7^3 Gauss points, abstract_basis.py:85-88) are extracted by *calling*
the reference (read-only, /root/reference) and shipped as a constant table.  Other rules fall
back to the tensor-product evaluation.

Run in the build container only:  python tools/gen_hex2_tables.py
"""
import sys
import numpy as np

sys.path.insert(0, "/root/reference")
from skfem.element import ElementHex2  # noqa: E402
from skfem.quadrature import get_quadrature  # noqa: E402

e = ElementHex2()
out = {}
for order in (12,):
    X, W = get_quadrature(e.refdom, order)
    phi = np.empty((27, X.shape[1]))
    dphi = np.empty((27, 3, X.shape[1]))
    for b in range(27):
        v, g = e.lbasis(X, b)
        phi[b], dphi[b] = v, g
    out["X_%d" % order], out["phi_%d" % order], out["dphi_%d" % order] = X, phi, dphi
for dst in ("scikit-fem_b200/skfem_b200/data/hex2_tables.npz",):
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})
