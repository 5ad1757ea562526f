"""``FacetBasis``: quadrature on mesh facets (boundary by default).

Mirrors skfem/assembly/basis/facet_basis.py:20-140 for affine meshes
(``MeshTri`` / ``MeshTet``) with the monomial elements.  The reference maps
the facet rule to global points ``x = G(X)``, pulls them back to each facet's
element ``Y = invF(x)`` and re-evaluates ``lbasis`` there, so every facet has
its own local points; ``skb_facet_geometry`` / ``skb_facet_basis``
(csrc/skb_facet.cu) do the same on the device, in the same operation order.
Forms over a FacetBasis go through the traced path of :mod:`skfem_b200.form`
with ``w.x``, ``w.h`` and ``w.n`` as device fields.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from . import _lib
from .basis import CellBasis, _torch, default_device
from .dofs import Dofs
from .mesh import OrientedBoundary
from .element import _MonomialElement
from .quadrature import get_quadrature

logger = logging.getLogger(__name__)

POLY_MAXT = 12  # terms per polynomial, csrc/skb_facet.cu


def poly_tables(elem):
    """Monomial tables of a scalar element for ``skb_facet_basis``:
    coef (nbs, 1+dim, POLY_MAXT) float64, expo (same) int32 with one exponent
    per byte, nterm (nbs, 1+dim) int32."""
    if not isinstance(elem, _MonomialElement):
        raise NotImplementedError("FacetBasis: {} has no monomial tables".format(
            type(elem).__name__))
    nbs, dim = len(elem._phi), elem.dim
    coef = np.zeros((nbs, 1 + dim, POLY_MAXT))
    expo = np.zeros((nbs, 1 + dim, POLY_MAXT), dtype=np.int32)
    nterm = np.zeros((nbs, 1 + dim), dtype=np.int32)
    for b in range(nbs):
        for c, terms in enumerate([elem._phi[b]] + list(elem._dphi[b])):
            if len(terms) > POLY_MAXT:
                raise ValueError("polynomial with more than {} terms".format(POLY_MAXT))
            nterm[b, c] = len(terms)
            for k, (cf, ex) in enumerate(terms):
                coef[b, c, k] = cf
                expo[b, c, k] = sum(int(e) << (8 * d) for d, e in enumerate(ex))
    return coef, expo, nterm


class FacetBasis(CellBasis):
    """For fields defined on the boundary (or any set of facets) of the domain.

    >>> fb = FacetBasis(MeshTri().refined(2), ElementTriP1())
    >>> BilinearForm(lambda u, v, w: u * v).assemble(fb)    # boundary mass matrix
    """
    _native_ok = False  # element-local kernels assume cell quadrature

    def __init__(self, mesh, elem, mapping=None, intorder=None, quadrature=None, facets=None,
                 dofs=None, side=0, disable_doflocs=False):
        if mesh.refdom is not elem.refdom:
            raise ValueError("Incompatible Mesh and Element.")
        if not mesh.affine:
            raise NotImplementedError("FacetBasis: affine meshes (MeshTri, MeshTet) only")
        logger.info("Initializing {}({}, {})".format(type(self).__name__, type(mesh).__name__,
                                                     type(elem).__name__))
        self.mesh = mesh
        self.elem = elem
        self.mapping = mesh._mapping() if mapping is None else mapping
        self.dofs = Dofs(mesh, elem) if dofs is None else dofs
        self.Nbfun = self.dofs.element_dofs.shape[0]
        self._tables = poly_tables(elem.scalar_element)
        if quadrature is not None:
            self.X, self.W = quadrature
        else:
            self.X, self.W = get_quadrature(mesh.refdom.brefdom, intorder if intorder is not None
                                            else 2 * elem.maxdeg)
        self.X = np.ascontiguousarray(self.X, dtype=np.float64)
        self.W = np.ascontiguousarray(self.W, dtype=np.float64)
        # by default use boundary facets (facet_basis.py:76-89)
        if facets is None:
            self.find = np.nonzero(mesh.f2t[1] == -1)[0].astype(np.int32)
        else:
            self.find = mesh.normalize_facets(facets)
        self._side = side
        self._facets_arg = self.find
        if isinstance(self.find, OrientedBoundary):
            # fix the orientation (facet_basis.py:84-87): traces from the oriented side (or,
            # side=1, from the other one), normals always from the oriented side
            ori = self.find.ori
            self.find = np.asarray(self.find)
            self.tind = mesh.f2t[(-1) ** side * ori - side, self.find]
            self.tind_normals = mesh.f2t[ori, self.find]
        else:
            self.find = np.asarray(self.find)
            self.tind = mesh.f2t[side, self.find]
            self.tind_normals = mesh.f2t[0, self.find]
        if len(self.find) == 0:
            logger.warning("Initializing {} with no facets.".format(type(self).__name__))
        elif self.tind.min() < 0:
            raise ValueError("side={}: some facets have no element on that side".format(side))
        self.nelems = len(self.find)
        self._affine = True
        self._disable_doflocs = disable_doflocs
        self._devcache = {}
        self._plans = {}
        self._fields = {}
        logger.info("Initializing finished.")

    def with_element(self, elem):
        """Same facets, side and quadrature with another element (facet_basis.py:257-270)."""
        return type(self)(self.mesh, elem, mapping=self.mapping, quadrature=(self.X, self.W),
                          facets=self._facets_arg, side=self._side)

    @property
    def nbs(self):
        return self._tables[0].shape[0]

    @property
    def element_dofs(self):
        if not hasattr(self, "_element_dofs"):
            self._element_dofs = np.ascontiguousarray(self.dofs.element_dofs[:, self.tind])
        return self._element_dofs

    # -- device residency -----------------------------------------------------------
    def _dev(self, device=None):
        torch = _torch()
        device = default_device() if device is None else device
        key = str(device)
        d = self._devcache.get(key)
        if d is not None:
            return d
        p, t = self.mesh.device_arrays(device)
        d = {"device": device, "p": p, "t": t}

        def up(a, dtype=None):
            a = np.ascontiguousarray(a if dtype is None else np.asarray(a).astype(dtype))
            return torch.from_numpy(a).to(device)
        m = self.mesh
        d["edofs"] = up(self.element_dofs)
        d["facets"] = up(m.facets, np.int32)
        d["find"], d["tind"] = up(self.find, np.int32), up(self.tind, np.int32)
        d["tind_n"] = up(self.tind_normals, np.int32)
        # local index of each facet inside the element the normal is taken from
        # (mapping_affine.py:266-269)
        hit = m.t2f[:, self.tind_normals] == np.asarray(self.find)[None, :]
        assert self.nelems == 0 or hit.any(axis=0).all()
        d["lfacet"] = up(np.argmax(hit, axis=0) if self.nelems else np.zeros(0), np.int32)
        d["X"], d["W"] = up(self.X), up(self.W)
        d["poly"] = tuple(up(a) for a in self._tables)
        sp = _lib.SkbSpace()
        sp.dim = m.dim()
        sp.nnodes = m.t.shape[0]
        sp.mapping = _lib.SKB_MAP_AFFINE
        sp.nbs = self.nbs
        sp.ncomp = self.ncomp
        sp.nqp = self.nqp
        sp.npts = m.p.shape[1]
        sp.nel_total = m.nelements
        sp.p, sp.t = p.data_ptr(), t.data_ptr()
        sp.tind = d["tind"].data_ptr()
        sp.nel = self.nelems
        sp.W = d["W"].data_ptr()
        d["space"] = sp
        self._devcache[key] = d
        return d

    def _tabulate(self, b=None, want=("grad",)):
        raise NotImplementedError("cell tabulation does not apply to a FacetBasis")

    def _geom(self):
        """x, Y, dx, n, detabs of all facets (one launch, cached)."""
        if "geom" not in self._fields:
            torch = _torch()
            d = self._dev()
            dev, nf, nqp, dim = d["device"], self.nelems, self.nqp, self.mesh.dim()
            g = {"x": torch.empty((dim, nf, nqp), dtype=torch.float64, device=dev),
                 "Y": torch.empty((dim, nf, nqp), dtype=torch.float64, device=dev),
                 "dx": torch.empty((nf, nqp), dtype=torch.float64, device=dev),
                 "n": torch.empty((dim, nf, nqp), dtype=torch.float64, device=dev),
                 "detabs": torch.empty((nf, nqp), dtype=torch.float64, device=dev)}
            code = _lib.lib().skb_facet_geometry(
                C.byref(d["space"]), d["facets"].data_ptr(), self.mesh.facets.shape[1],
                d["find"].data_ptr(), d["tind"].data_ptr(), d["tind_n"].data_ptr(),
                d["lfacet"].data_ptr(), nf, d["X"].data_ptr(), d["W"].data_ptr(), nqp,
                g["x"].data_ptr(), g["Y"].data_ptr(), g["dx"].data_ptr(), g["n"].data_ptr(),
                g["detabs"].data_ptr(), self._stream())
            _lib.check(code, "skb_facet_geometry")
            self._fields["geom"] = g
        return self._fields["geom"]

    def _dx_dev(self):
        return self._geom()["dx"]

    def _scalar_basis_dev(self, b):
        key = ("basis", b)
        if key not in self._fields:
            torch = _torch()
            d = self._dev()
            dev, nf, nqp, dim = d["device"], self.nelems, self.nqp, self.mesh.dim()
            val = torch.empty((nf, nqp), dtype=torch.float64, device=dev)
            grad = torch.empty((dim, nf, nqp), dtype=torch.float64, device=dev)
            coef, expo, nterm = d["poly"]
            code = _lib.lib().skb_facet_basis(
                C.byref(d["space"]), d["tind"].data_ptr(), nf, nqp, self._geom()["Y"].data_ptr(),
                coef.data_ptr(), expo.data_ptr(), nterm.data_ptr(), int(b), val.data_ptr(),
                grad.data_ptr(), self._stream())
            _lib.check(code, "skb_facet_basis")
            self._fields[key] = (val, grad)
        return self._fields[key]

    def _basis_field_dev(self, i):
        """Global basis function ``i`` at the facet quadrature points
        (facet_basis.py:111-112 -> ElementH1.gbasis / ElementVector.gbasis)."""
        from .field import DiscreteField
        torch = _torch()
        nc = self.ncomp
        b, n = divmod(i, nc)
        val, g = self._scalar_basis_dev(b)
        if nc == 1:
            return DiscreteField(val, g)
        dim, nf, nqp = self.mesh.dim(), self.nelems, self.nqp
        vv = torch.zeros((dim, nf, nqp), dtype=torch.float64, device=val.device)
        gg = torch.zeros((dim, dim, nf, nqp), dtype=torch.float64, device=val.device)
        vv[n] = val
        gg[n] = g
        return DiscreteField(vv, gg)

    @property
    def normals(self):
        from .field import DiscreteField
        return DiscreteField(self._geom()["n"])

    def global_coordinates(self):
        from .field import DiscreteField
        return DiscreteField(self._geom()["x"])

    def mesh_parameters(self):
        """|detDG| ** (1 / (dim - 1)) (facet_basis.py:133-140)."""
        from .field import DiscreteField
        if "h" not in self._fields:
            self._fields["h"] = self._geom()["detabs"] ** (1. / (self.mesh.dim() - 1.))
        return DiscreteField(self._fields["h"], lay="F")

    @property
    def _dx_layout(self):
        # |tile(detB, (nqp, 1)).T| * broadcast W: Fortran-ordered (mapping_affine.py:242-246)
        return "F"

    def default_parameters(self):
        return {"x": self.global_coordinates(), "h": self.mesh_parameters(), "n": self.normals}


BoundaryFacetBasis = FacetBasis  # deprecated alias kept by the reference


class InteriorFacetBasis(FacetBasis):
    """FacetBasis over the interior facets by default; ``side`` selects which of
    the two adjacent elements is traced
    (skfem/assembly/basis/interior_facet_basis.py:12-50)."""

    def __init__(self, mesh, elem, mapping=None, intorder=None, quadrature=None, facets=None,
                 dofs=None, side=0, disable_doflocs=False):
        if facets is None:
            facets = np.nonzero(mesh.f2t[1] != -1)[0].astype(np.int32)
        super().__init__(mesh, elem, mapping=mapping, intorder=intorder, quadrature=quadrature,
                         facets=facets, dofs=dofs, side=side, disable_doflocs=disable_doflocs)
