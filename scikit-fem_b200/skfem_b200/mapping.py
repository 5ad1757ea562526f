"""Reference-to-global mappings.

In the reference these classes own the geometry arrays (``A, invA, detA`` of
skfem/mapping/mapping_affine.py:55-131, the cached Jacobians of
mapping_isoparametric.py:112-135).  In this engine geometry never exists as
arrays on the hot path: each kernel recomputes ``F, DF, invDF, detDF`` per
element (per quadrature point for hexahedra) in registers from coalesced loads
of ``p`` and ``t``.  The classes remain as the API objects a ``CellBasis`` is
parameterised with; their array-returning methods (``F``, ``DF``, ``invDF``, ``detDF``,
``invF`` - the Mapping contract of skfem/mapping/mapping.py:6-114) evaluate on the device
on demand (``skb_tabulate`` / ``skb_mapping``) and copy back.
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

    def _mapping_arrays(self, X, tind, want):
        """DF / invDF (dim, dim, nel, npts) and the signed detDF (nel, npts) from the device
        (``skb_mapping``), host numpy like the reference's return values."""
        import ctypes as C
        import torch
        from . import _lib
        b = self._basis(X, tind)
        d = b._dev()
        dim, nel, nqp = self.dim, b.nelems, b.nqp
        out = {}
        if "DF" in want:
            out["DF"] = torch.empty((dim, dim, nel, nqp), dtype=torch.float64, device=d["device"])
        if "invDF" in want:
            out["invDF"] = torch.empty((dim, dim, nel, nqp), dtype=torch.float64,
                                       device=d["device"])
        if "det" in want:
            out["det"] = torch.empty((nel, nqp), dtype=torch.float64, device=d["device"])

        def ptr(k):
            return out[k].data_ptr() if k in out else None
        code = _lib.lib().skb_mapping(C.byref(d["space"]), ptr("DF"), ptr("invDF"), ptr("det"),
                                      b._stream())
        _lib.check(code, "skb_mapping")
        return {k: v.cpu().numpy() for k, v in out.items()}

    def detDF(self, X, tind=None):
        """Signed Jacobian determinant (nel, npts) (mapping_affine.py:205-211,
        mapping_isoparametric.py:179-198; the latter raises on a zero determinant)."""
        return self._mapping_arrays(X, tind, ("det",))["det"]

    def DF(self, X, tind=None):
        """Jacobian (dim, dim, nel, npts) (mapping_affine.py:213-223 /
        mapping_isoparametric.py:173-177)."""
        return self._mapping_arrays(X, tind, ("DF",))["DF"]

    def invDF(self, X, tind=None):
        """Inverse Jacobian (dim, dim, nel, npts) (mapping_affine.py:225-232 /
        mapping_isoparametric.py:200-226)."""
        return self._mapping_arrays(X, tind, ("invDF",))["invDF"]

    def invF(self, x, tind=None):
        """Local coordinates of global points ``x`` (dim, nel, npts): the affine inverse map
        ``invA (x - b)`` (mapping_affine.py:195-203), the products of each row summed in
        column order like the reference's einsum.  The Newton iteration of the isoparametric
        map (mapping_isoparametric.py:157-168) is not on the assembly path."""
        if not isinstance(self, MappingAffine):
            raise NotImplementedError("invF of an isoparametric mapping")
        import torch
        X0 = np.zeros((self.dim, 1))
        b = self._basis(X0, tind)
        dev = b._dev()["device"]
        inv = torch.as_tensor(self._mapping_arrays(X0, tind, ("invDF",))["invDF"][..., 0],
                              device=dev)                       # (dim, dim, nel)
        orig = b.global_coordinates().t[:, :, 0]                # F(0) = b: (dim, nel)
        y = torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev) - orig[:, :, None]
        out = []
        for i in range(self.dim):
            acc = inv[i, 0][:, None] * y[0]
            for j in range(1, self.dim):
                acc = acc + inv[i, j][:, None] * y[j]
            out.append(acc)
        return torch.stack(out).cpu().numpy()


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
