"""Reference-to-global mappings.

In the reference these classes own the geometry arrays (``A, invA, detA`` of
skfem/mapping/mapping_affine.py:55-131, the cached Jacobians of
mapping_isoparametric.py:112-135).  In this engine geometry never exists as
arrays on the hot path: each kernel recomputes ``F, DF, invDF, detDF`` per
element (per quadrature point for hexahedra) in registers from coalesced loads
of ``p`` and ``t``.  The classes remain as the API objects a ``CellBasis`` is
parameterised with; their array-returning methods evaluate on the device on
demand (``skb_tabulate``) and copy back.
"""
import weakref

import numpy as np


class Mapping:
    def __init__(self, mesh):
        # weak: the mesh caches its mapping (Mesh._mapping); a strong back-reference would
        # make every mesh (and the device copies of p and t it owns) wait for the cyclic GC
        self._mesh = weakref.ref(mesh)
        self.dim = mesh.p.shape[0]

    @property
    def mesh(self):
        m = self._mesh()
        if m is None:
            raise ReferenceError("the mesh of this mapping no longer exists")
        return m

    def _basis(self, X, tind):
        from .basis import CellBasis
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise NotImplementedError("per-element local points are not supported")
        return CellBasis(self.mesh, self.mesh.elem(), quadrature=(X, np.ones(X.shape[1])),
                         elements=tind, disable_doflocs=True)

    def F(self, X, tind=None):
        """Global coordinates of local points X: (dim, nel, npts)."""
        return self._basis(X, tind).global_coordinates().numpy()

    def detDF(self, X, tind=None):
        """Signed Jacobian determinants are not exposed by the kernels (only
        their absolute value enters ``dx``); returns |detDF| (nel, npts)."""
        b = self._basis(X, tind)
        return b._tabulate(want=("detabs",))["detabs"].cpu().numpy()


class MappingAffine(Mapping):
    """Affine map of simplices (tri, tet)."""

    def __init__(self, mesh, tind=None):
        super().__init__(mesh)
        self.tind = tind


class MappingIsoparametric(Mapping):
    """Isoparametric map defined by the mesh's geometry element (Hex1)."""

    def __init__(self, mesh, elem, bndelem=None):
        super().__init__(mesh)
        self.elem = elem
        self.bndelem = bndelem
