"""Finite elements of the hot path: H1-conforming Lagrange elements defined on a
reference domain, plus the vector-valued wrapper.

Mirrors the reference's ``Element`` interface (skfem/element/element.py:12-99):
class attributes ``nodal_dofs / edge_dofs / facet_dofs / interior_dofs /
maxdeg / refdom / doflocs`` and ``lbasis(X, i) -> (phi, dphi)``.

Only ``lbasis`` runs on the host: the (nbs, nqp) value table and the
(nbs, dim, nqp) reference-gradient table are uploaded once and the push-forward
through ``invDF`` happens inside the CUDA kernels (the reference's
``ElementH1.gbasis``, skfem/element/element_h1.py:10-24, materialises
(dim, nel, nqp) arrays per basis function instead).

Simplex bases are stored as ordered monomial lists.  The term order equals the
order in which the reference writes its polynomials
(element_tri/element_tri_p1.py:18-33, element_tri_p2.py:22-50,
element_tet/element_tet_p1.py:19-45, element_tet_p2.py:26-103), which makes the
tabulated values bit-identical (every multi-variable coefficient is a power of
two, so the association of the products cannot change the rounding).
"""
from __future__ import annotations

import numpy as np

from .refdom import Refdom, RefTri, RefTet, RefHex, RefQuad, RefLine


class Element:
    nodal_dofs = 0
    facet_dofs = 0
    interior_dofs = 0
    edge_dofs = 0
    maxdeg = -1
    dofnames: list = []
    refdom = Refdom
    doflocs: np.ndarray

    @property
    def dim(self):
        return self.refdom.dim()

    def __call__(self):
        return self

    def lbasis(self, X, i):
        raise NotImplementedError

    @classmethod
    def _index_error(cls):
        raise ValueError("Index larger than the number of basis functions.")

    def _bfun_counts(self):
        rd = self.refdom
        return np.array([self.nodal_dofs * rd.nnodes, self.edge_dofs * rd.nedges,
                         self.facet_dofs * rd.nfacets, self.interior_dofs])

    @property
    def nbfun(self):
        """Local basis functions per element (incl. vector components)."""
        rd = self.refdom
        n = self.nodal_dofs * rd.nnodes + self.facet_dofs * rd.nfacets + self.interior_dofs
        if rd.dim() == 3:
            n += self.edge_dofs * rd.nedges
        return int(n)

    def gbasis(self, mapping, X, i, tind=None):
        """Global basis function ``i`` at the local points ``X`` (dim, npts) of the elements
        ``tind`` - the reference's ``Element.gbasis(mapping, X, i, tind)``
        (element/element.py:67-99, element_h1.py:10-24, element_vector.py:36-48): a 1-tuple
        holding a DiscreteField with ``value (nel, npts)`` and ``grad (dim, nel, npts)``
        (vector elements: ``(dim, nel, npts)`` / ``(dim, dim, nel, npts)``), here resident on
        the device (``.numpy()`` / ``np.asarray`` bring it to the host).  The push-forward runs
        in ``skb_tabulate``; the ``(dim, dim, nel, npts)`` inverse Jacobians are never formed."""
        from .basis import CellBasis
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise NotImplementedError("per-element local points are not supported")
        if not 0 <= i < self.nbfun:
            self._index_error()
        b = CellBasis(mapping.mesh, self, mapping=mapping,
                      quadrature=(X, np.ones(X.shape[1])), elements=tind, disable_doflocs=True)
        return (b._basis_field_dev(i),)

    # -- host tables consumed by the kernels --------------------------------
    def tabulate(self, X):
        """(phi (nbs, nqp), dphi (nbs, dim, nqp)) of the *scalar* basis."""
        nbs = self.scalar_element.nbfun
        nqp = X.shape[1]
        phi = np.empty((nbs, nqp))
        dphi = np.empty((nbs, X.shape[0], nqp))
        for b in range(nbs):
            v, g = self.scalar_element.lbasis(X, b)
            phi[b] = v
            dphi[b] = g
        return phi, dphi

    @property
    def scalar_element(self):
        return self

    @property
    def ncomp(self):
        return 1


class ElementH1(Element):
    """H1-conforming element: identity push-forward of values, gradients
    through invDF^T (done on the GPU)."""


# ---------------------------------------------------------------------------
# simplex Lagrange elements as ordered monomial lists
# term = (coefficient, (ex, ey[, ez]))
# ---------------------------------------------------------------------------
def _mono(X, expo):
    out = None
    for v, k in zip(X, expo):
        for _ in range(k):
            out = v if out is None else out * v
    return out


def _poly(X, terms):
    acc = None
    for c, expo in terms:
        m = _mono(X, expo)
        t = c if m is None else c * m
        acc = t if acc is None else acc + t
    if np.ndim(acc) == 0:  # constant polynomial -> broadcast like ``c + 0*x``
        acc = acc + 0 * X[0]
    return acc


class _MonomialElement(ElementH1):
    _phi: list = []
    _dphi: list = []

    def lbasis(self, X, i):
        if not 0 <= i < len(self._phi):
            self._index_error()
        return _poly(X, self._phi[i]), np.array([_poly(X, d) for d in self._dphi[i]])


_Z2, _X2, _Y2 = (0, 0), (1, 0), (0, 1)
_Z3, _X3, _Y3, _ZZ3 = (0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)


class ElementTriP1(_MonomialElement):
    """Piecewise linear triangle."""
    nodal_dofs = 1
    maxdeg = 1
    dofnames = ['u']
    refdom = RefTri
    doflocs = np.array([[0., 0.], [1., 0.], [0., 1.]])
    _phi = [[(1., _Z2), (-1., _X2), (-1., _Y2)], [(1., _X2)], [(1., _Y2)]]
    _dphi = [[[(-1., _Z2), (0., _X2)], [(-1., _Z2), (0., _X2)]],
             [[(1., _Z2), (0., _X2)], [(0., _X2)]],
             [[(0., _X2)], [(1., _Z2), (0., _X2)]]]


class ElementTriP2(_MonomialElement):
    """Piecewise quadratic triangle."""
    nodal_dofs = 1
    facet_dofs = 1
    maxdeg = 2
    dofnames = ['u', 'u']
    refdom = RefTri
    doflocs = np.array([[0., 0.], [1., 0.], [0., 1.], [.5, 0.], [.5, .5], [0., .5]])
    _g0 = [(-3., _Z2), (4., _X2), (4., _Y2)]
    _phi = [
        [(1., _Z2), (-3., _X2), (-3., _Y2), (2., (2, 0)), (4., (1, 1)), (2., (0, 2))],
        [(2., (2, 0)), (-1., _X2)],
        [(2., (0, 2)), (-1., _Y2)],
        [(4., _X2), (-4., (2, 0)), (-4., (1, 1))],
        [(4., (1, 1))],
        [(4., _Y2), (-4., (1, 1)), (-4., (0, 2))],
    ]
    _dphi = [
        [_g0, _g0],
        [[(4., _X2), (-1., _Z2)], [(0., _X2)]],
        [[(0., _X2)], [(4., _Y2), (-1., _Z2)]],
        [[(4., _Z2), (-8., _X2), (-4., _Y2)], [(-4., _X2)]],
        [[(4., _Y2)], [(4., _X2)]],
        [[(-4., _Y2)], [(4., _Z2), (-4., _X2), (-8., _Y2)]],
    ]


class ElementTetP1(_MonomialElement):
    """Piecewise linear tetrahedron."""
    nodal_dofs = 1
    maxdeg = 1
    dofnames = ['u']
    refdom = RefTet
    doflocs = np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
    _one = [(1., _Z3), (0., _X3)]
    _nil = [(0., _X3)]
    _neg = [(-1., _Z3), (0., _X3)]
    _phi = [[(1., _Z3), (-1., _X3), (-1., _Y3), (-1., _ZZ3)],
            [(1., _X3)], [(1., _Y3)], [(1., _ZZ3)]]
    _dphi = [[_neg, _neg, _neg], [_one, _nil, _nil], [_nil, _one, _nil], [_nil, _nil, _one]]


class ElementTetP2(_MonomialElement):
    """Piecewise quadratic tetrahedron (vertex + edge-midpoint DOFs)."""
    nodal_dofs = 1
    edge_dofs = 1
    maxdeg = 2
    dofnames = ['u', 'u']
    refdom = RefTet
    doflocs = np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.],
                        [.5, 0., 0.], [.5, .5, 0.], [0., .5, 0.],
                        [0., 0., .5], [.5, 0., .5], [0., .5, .5]])
    _nil = [(0., _X3)]
    _g0 = [(-3., _Z3), (4., _X3), (4., _Y3), (4., _ZZ3)]
    _phi = [
        [(1., _Z3), (-3., _X3), (2., (2, 0, 0)), (-3., _Y3), (4., (1, 1, 0)),
         (2., (0, 2, 0)), (-3., _ZZ3), (4., (1, 0, 1)), (4., (0, 1, 1)), (2., (0, 0, 2))],
        [(-1., _X3), (2., (2, 0, 0))],
        [(-1., _Y3), (2., (0, 2, 0))],
        [(-1., _ZZ3), (2., (0, 0, 2))],
        [(4., _X3), (-4., (2, 0, 0)), (-4., (1, 1, 0)), (-4., (1, 0, 1))],
        [(4., (1, 1, 0))],
        [(0., _Z3), (4., _Y3), (-4., (1, 1, 0)), (-4., (0, 2, 0)), (-4., (0, 1, 1))],
        [(0., _Z3), (4., _ZZ3), (-4., (1, 0, 1)), (-4., (0, 1, 1)), (-4., (0, 0, 2))],
        [(0., _Z3), (4., (1, 0, 1))],
        [(0., _Z3), (4., (0, 1, 1))],
    ]
    _dphi = [
        [_g0, _g0, _g0],
        [[(-1., _Z3), (4., _X3)], _nil, _nil],
        [_nil, [(-1., _Z3), (4., _Y3)], _nil],
        [_nil, _nil, [(-1., _Z3), (4., _ZZ3)]],
        [[(4., _Z3), (-8., _X3), (-4., _Y3), (-4., _ZZ3)], [(-4., _X3)], [(-4., _X3)]],
        [[(4., _Y3)], [(4., _X3)], _nil],
        [[(-4., _Y3)], [(4., _Z3), (-4., _X3), (-8., _Y3), (-4., _ZZ3)], [(-4., _Y3)]],
        [[(-4., _ZZ3)], [(-4., _ZZ3)], [(4., _Z3), (-4., _X3), (-4., _Y3), (-8., _ZZ3)]],
        [[(4., _ZZ3)], _nil, [(4., _X3)]],
        [_nil, [(4., _ZZ3)], [(4., _Y3)]],
    ]


# ---------------------------------------------------------------------------
# hexahedra: tensor products of 1-D Lagrange bases on the RefHex numbering
# ---------------------------------------------------------------------------
class ElementHex1(ElementH1):
    """Trilinear hexahedron; also the geometry element of ``MeshHex``.

    Basis k belongs to reference vertex ``RefHex.p[:, k]``: the product of
    ``c`` (vertex coordinate 1) or ``1 - c`` (vertex coordinate 0) over the
    three axes, multiplied left to right like
    skfem/element/element_hex/element_hex1.py:23-69."""
    nodal_dofs = 1
    maxdeg = 3
    dofnames = ['u']
    refdom = RefHex
    doflocs = RefHex.p.T.copy()

    def lbasis(self, X, i):
        if not 0 <= i < 8:
            self._index_error()
        corner = RefHex.p[:, i]
        f = [c if k == 1. else 1 - c for c, k in zip(X, corner)]
        sgn = [1. if k == 1. else -1. for k in corner]
        phi = f[0] * f[1] * f[2]
        pairs = ((1, 2), (0, 2), (0, 1))
        dphi = np.array([f[a] * f[b] if s > 0 else -f[a] * f[b]
                         for s, (a, b) in zip(sgn, pairs)])
        return phi, dphi


def _lagrange2(node, x):
    """1-D quadratic Lagrange basis on nodes {0, 1/2, 1}: value, derivative."""
    if node == 0.:
        return (2. * x - 1.) * (x - 1.), 4. * x - 3.
    if node == 1.:
        return x * (2. * x - 1.), 4. * x - 1.
    return 4. * x * (1. - x), 4. - 8. * x


class ElementHex2(ElementH1):
    """Triquadratic hexahedron (27 nodes: vertices, edge / facet / cell
    centres in RefHex order, cf. skfem/element/element_hex/element_hex2.py:
    1213-1260).  The reference evaluates machine-generated Horner forms.  At the default
    quadrature rule (7^3 Gauss points) ``tabulate`` returns the reference's own numbers from a
    shipped constant table (data/hex2_tables.npz, tools/gen_hex2_tables.py); at
    any other points the tensor-product evaluation below, which agrees with the reference to
    a few ulp (tests/test_host_api.py), is used."""
    _tables = None

    def tabulate(self, X):
        import os
        if ElementHex2._tables is None:
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data",
                                "hex2_tables.npz")
            ElementHex2._tables = dict(np.load(path)) if os.path.exists(path) else {}
        for key, Xt in ElementHex2._tables.items():
            if key.startswith("X_") and Xt.shape == X.shape and np.array_equal(Xt, X):
                order = key[2:]
                return (ElementHex2._tables["phi_" + order].copy(),
                        ElementHex2._tables["dphi_" + order].copy())
        return super().tabulate(X)

    nodal_dofs = 1
    facet_dofs = 1
    edge_dofs = 1
    interior_dofs = 1
    maxdeg = 6
    dofnames = ['u', 'u', 'u', 'u']
    refdom = RefHex
    doflocs = np.vstack([
        RefHex.p.T,
        [RefHex.p[:, e].mean(axis=1) for e in RefHex.edges],
        [RefHex.p[:, f].mean(axis=1) for f in RefHex.facets],
        [[.5, .5, .5]],
    ])

    def lbasis(self, X, i):
        if not 0 <= i < 27:
            self._index_error()
        (a, da), (b, db), (c, dc) = (_lagrange2(n, x) for n, x in zip(self.doflocs[i], X))
        return a * b * c, np.array([da * b * c, a * db * c, a * b * dc])


class ElementVector(Element):
    """The same scalar element for every vector component; local basis
    function ``i`` is scalar function ``i // dim`` in component ``i % dim``
    (skfem/element/element_vector.py:8-48)."""

    def __init__(self, elem, dim=None):
        self.elem = elem
        self._dim = elem.dim if dim is None else dim
        if self._dim != elem.dim:
            raise NotImplementedError("ElementVector: dim must equal the spatial dimension")
        self.nodal_dofs = elem.nodal_dofs * self._dim
        self.facet_dofs = elem.facet_dofs * self._dim
        self.interior_dofs = elem.interior_dofs * self._dim
        self.edge_dofs = elem.edge_dofs * self._dim
        self.dofnames = [n + "^" + str(j + 1) for n in elem.dofnames for j in range(self._dim)]
        self.maxdeg = elem.maxdeg
        self.refdom = elem.refdom
        if hasattr(elem, 'doflocs'):
            self.doflocs = np.repeat(elem.doflocs, self._dim, axis=0)

    @property
    def dim(self):
        return self._dim

    @property
    def scalar_element(self):
        return self.elem

    @property
    def ncomp(self):
        return self._dim

    def lbasis(self, X, i):
        raise NotImplementedError("ElementVector has no scalar lbasis; "
                                  "use scalar_element.lbasis")


__all__ = ["Element", "ElementH1", "ElementTriP1", "ElementTriP2", "ElementTetP1",
           "ElementTetP2", "ElementHex1", "ElementHex2", "ElementVector",
           "RefTri", "RefTet", "RefHex", "RefQuad", "RefLine"]
