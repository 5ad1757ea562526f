"""Dump the simplex quadrature rules (points, weights) used on the hot path.

The rules are published numerical tables (Keast / Dunavant-type rules); the
reference keeps them in skfem/quadrature.py:80-2836.  To guarantee that the
B200 engine integrates with bit-identical (X, W) we extract the numbers by
*calling* the reference (read-only, /root/reference) and store them as a
binary table.  Tensor-product (hex/quad/line) rules are NOT stored: both the
product and the oracle regenerate them from numpy's leggauss exactly like
skfem/quadrature.py:55-74,2839-2844.

Run in the build container only:  python tools/gen_quadrature_tables.py
"""
import sys
import numpy as np

sys.path.insert(0, "/root/reference")
from skfem.quadrature import get_quadrature_tet, get_quadrature_tri  # noqa

out = {}
for name, fn in (("tet", get_quadrature_tet), ("tri", get_quadrature_tri)):
    orders = []
    for order in range(1, 40):
        try:
            X, W = fn(order)
        except NotImplementedError:
            continue
        out[f"{name}_{order}_X"] = np.ascontiguousarray(X, dtype=np.float64)
        out[f"{name}_{order}_W"] = np.ascontiguousarray(W, dtype=np.float64)
        orders.append(order)
    out[f"{name}_orders"] = np.array(orders, dtype=np.int64)

for dst in ("scikit-fem_b200/skfem_b200/data/quadrature_tables.npz",
            "oracle/quadrature_tables.npz"):
    np.savez_compressed(dst, **out)
    print("wrote", dst, len(out), "arrays")
