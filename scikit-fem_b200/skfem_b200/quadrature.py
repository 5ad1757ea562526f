"""Quadrature rules, bit-identical to the reference's (skfem/quadrature.py).

Simplex rules are published tables; they are shipped as data
(data/quadrature_tables.npz, extracted by tools/gen_quadrature_tables.py).
Line / quad / hex rules are generated from numpy's Gauss-Legendre nodes with
the same arithmetic as skfem/quadrature.py:55-74,2839-2844 (meshgrid + Fortran
flattening decides the point order, which decides the summation order).
"""
import os
from functools import lru_cache

import numpy as np
from numpy.polynomial.legendre import leggauss

from .refdom import RefLine, RefTri, RefTet, RefQuad, RefHex

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data",
                     "quadrature_tables.npz")


@lru_cache(maxsize=None)
def _tables():
    with np.load(_DATA) as z:
        return {k: z[k] for k in z.files}


def _simplex(kind, order):
    order = max(int(order), 1)
    tab = _tables()
    try:
        return tab[f"{kind}_{order}_X"].copy(), tab[f"{kind}_{order}_W"].copy()
    except KeyError:
        raise NotImplementedError("The requested order of quadrature"
                                  "is not implemented!")


def _line(order):
    if order <= 1:
        order = 2
    x, w = leggauss(int(np.ceil((order + 1.0) / 2.0)))
    return np.array([0.5 * x + 0.5]), w / 2.0


def _tensor(order, dim):
    x1, w1 = _line(order)
    pts = np.meshgrid(*(dim * (x1,)))
    wts = np.meshgrid(*(dim * (w1,)))
    X = np.vstack([a.flatten(order="F") for a in pts])
    W = wts[0]
    for a in wts[1:]:
        W = W * a
    return X, W.flatten(order="F")


def get_quadrature(refdom_or_elem, norder):
    """(X (dim, nqp), W (nqp,)) exact to polynomial degree ``norder``."""
    refdom = getattr(refdom_or_elem, "refdom", refdom_or_elem)
    if refdom is RefTri:
        return _simplex("tri", norder)
    if refdom is RefTet:
        return _simplex("tet", norder)
    if refdom is RefLine:
        return _line(norder)
    if refdom is RefQuad:
        return _tensor(norder, 2)
    if refdom is RefHex:
        return _tensor(norder, 3)
    raise NotImplementedError("The given reference domain type '{}' "
                              "is not supported!".format(refdom))
