"""``CellBasis``: a mesh + an element + a quadrature rule, resident on the GPU.

The reference's constructor (skfem/assembly/basis/cell_basis.py:42-106 on top
of abstract_basis.py:45-88) eagerly materialises, for every local basis
function, a ``(dim, nel, nqp)`` gradient array and the ``(nel, nqp)`` ``dx``
array.  Here the constructor only

* numbers the DOFs (host, :mod:`skfem_b200.dofs`),
* picks the quadrature rule and tabulates ``lbasis`` at its points (host,
  a few kB),
* uploads ``p``, ``t``, ``element_dofs`` and the tables.

Geometry and push-forward are recomputed inside each assembly kernel from ``p``
and ``t``.  ``basis.dx`` / ``basis.basis`` / ``basis.default_parameters()``
stay available with the reference's shapes - they are produced on demand by
``skb_tabulate`` (and cached) because traced user forms need them.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from . import _lib
from .dofs import Dofs
from .quadrature import get_quadrature

logger = logging.getLogger(__name__)


def _torch():
    import torch
    return torch


def default_device():
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("skfem_b200 needs a CUDA device (B200); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


class AbstractBasis:
    """Attributes shared by all bases (abstract_basis.py:20-44)."""
    tind = None


class CellBasis(AbstractBasis):

    def __init__(self, mesh, elem, mapping=None, intorder=None, elements=None,
                 quadrature=None, dofs=None, disable_doflocs=False):
        if mesh.refdom is not elem.refdom:
            raise ValueError("Incompatible Mesh and Element.")
        logger.info("Initializing {}({}, {})".format(type(self).__name__, type(mesh).__name__,
                                                     type(elem).__name__))
        self.mesh = mesh
        self.elem = elem
        self.mapping = mesh._mapping() if mapping is None else mapping
        self.dofs = Dofs(mesh, elem) if dofs is None else dofs
        self.Nbfun = self.dofs.element_dofs.shape[0]
        if quadrature is not None:
            self.X, self.W = quadrature
        else:
            self.X, self.W = get_quadrature(mesh.refdom, intorder if intorder is not None
                                            else 2 * elem.maxdeg)
        self.X = np.ascontiguousarray(self.X, dtype=np.float64)
        self.W = np.ascontiguousarray(self.W, dtype=np.float64)
        if elements is None:
            self.tind = None
            self.nelems = mesh.nelements
        else:
            self.tind = mesh.normalize_elements(elements)
            self.nelems = len(self.tind)
        # reference tables (host): scalar element + geometry element
        self._phi, self._dphi = elem.tabulate(self.X)
        self._affine = bool(mesh.affine)
        if not self._affine:
            self._mphi, self._mdphi = mesh.elem().tabulate(self.X)
        self._disable_doflocs = disable_doflocs
        self._devcache = {}
        self._plans = {}
        self._fields = {}
        logger.info("Initializing finished.")

    # -- reference attributes ------------------------------------------------------
    @property
    def N(self):
        return self.dofs.N

    @property
    def nodal_dofs(self):
        return self.dofs.nodal_dofs

    @property
    def edge_dofs(self):
        return self.dofs.edge_dofs

    @property
    def facet_dofs(self):
        return self.dofs.facet_dofs

    @property
    def interior_dofs(self):
        return self.dofs.interior_dofs

    @property
    def element_dofs(self):
        if not hasattr(self, "_element_dofs"):
            ed = self.dofs.element_dofs
            self._element_dofs = ed if self.tind is None else np.ascontiguousarray(ed[:, self.tind])
        return self._element_dofs

    @property
    def nqp(self):
        return self.W.shape[-1]

    @property
    def nbs(self):
        return self._phi.shape[0]

    @property
    def ncomp(self):
        return self.elem.ncomp

    def get_dofs(self, facets=None, elements=None, nodes=None, skip=None):
        """DOFs on a set of facets (abstract_basis.py:124-237): ``facets`` is None (the
        whole boundary), an array of facet indices, a callable on facet midpoints, a
        boundary name or a list of those; a dict of names gives a dict of results.
        ``elements`` / ``nodes`` select by elements (anything ``normalize_elements`` takes) or
        vertices instead; ``skip`` lists DOF names to leave out.  Returns a
        :class:`~skfem_b200.dofs.DofsView`: array-like (sorted unique int32 indices) with
        ``.all(name)``, ``.flatten()``, ``.keep / .drop``, ``.nodal / .facet / .edge /
        .interior``."""
        locs = (lambda: self.doflocs) if not self._disable_doflocs else None
        if isinstance(facets, dict):
            return {k: self.get_dofs(v, skip=skip) for k, v in facets.items()}
        if elements is not None:
            return self.dofs.element_view(self.mesh.normalize_elements(elements), skip, locs)
        if nodes is not None:
            return self.dofs.vertex_view(self.mesh.normalize_nodes(nodes), skip, locs)
        facets = (self.mesh.boundary_facets() if facets is None
                  else self.mesh.normalize_facets(facets))
        return self.dofs.facet_view(facets, skip, locs)

    def with_element(self, elem):
        """Same mesh, quadrature and element subset with another element
        (cell_basis.py:262-275)."""
        return type(self)(self.mesh, elem, mapping=self.mapping, quadrature=(self.X, self.W),
                          elements=self.tind)

    def with_elements(self, elements):
        """Same element and quadrature restricted to ``elements`` - anything
        ``Mesh.normalize_elements`` accepts (cell_basis.py:266-280)."""
        return type(self)(self.mesh, self.elem, mapping=self.mapping,
                          quadrature=(self.X, self.W), elements=elements)

    def boundary(self, facets=None, intorder=None, quadrature=None):
        """The FacetBasis of the same mesh and element (cell_basis.py:282-308)."""
        from .facet_basis import FacetBasis
        if self.tind is not None:
            raise NotImplementedError("Boundary of subdomain not supported.")
        return FacetBasis(self.mesh, self.elem, mapping=self.mapping, facets=facets,
                          intorder=intorder, quadrature=quadrature)

    @property
    def quadrature(self):
        return self.X, self.W

    def zero_w(self, dtype=None):
        """Zero array of the shape forms see at the quadrature points
        (abstract_basis.py:378-382)."""
        return np.zeros((self.nelems, len(self.W)), dtype=dtype)

    def complement_dofs(self, *D):
        return np.setdiff1d(np.arange(self.N), np.concatenate([np.asarray(d).ravel() for d in D]))

    def zeros(self):
        return np.zeros(self.N)

    def ones(self):
        return np.ones(self.N)

    @property
    def doflocs(self):
        """Global DOF locations (abstract_basis.py:62-73), host numpy."""
        if self._disable_doflocs:
            raise AttributeError("doflocs disabled")
        if not hasattr(self, "_doflocs"):
            Xd = np.ascontiguousarray(self.elem.doflocs.T)
            other = CellBasis(self.mesh, self.elem, quadrature=(Xd, np.ones(Xd.shape[1])),
                              dofs=self.dofs, disable_doflocs=True)
            x = other.global_coordinates().numpy()       # (dim, nel, Nbfun)
            out = np.zeros((x.shape[0], self.N))
            ed = self.dofs.element_dofs
            for i in range(x.shape[0]):
                for j in range(ed.shape[0]):
                    out[i, ed[j]] = x[i, :, j]
            self._doflocs = out
        return self._doflocs

    # -- device residency ---------------------------------------------------------------
    def _dev(self, device=None):
        """Upload (once) everything the kernels read; returns a dict of torch
        tensors plus the ctypes ``skb_space_t`` describing this basis."""
        torch = _torch()
        device = default_device() if device is None else device
        key = str(device)
        d = self._devcache.get(key)
        if d is not None:
            return d
        with _lib.nvtx("skfem_b200:basis-upload"):
            p, t = self.mesh.device_arrays(device)
        d = {"device": device, "p": p, "t": t}

        def up(a, dtype=None):
            a = np.ascontiguousarray(a)
            return torch.from_numpy(a).to(device)
        kept = getattr(self.dofs, "_edofs_dev", None)
        if self.element_dofs is self.mesh.t:
            d["edofs"] = t
        elif kept is not None and kept[0] == key and self.element_dofs is self.dofs.element_dofs:
            d["edofs"] = kept[1]                     # built on the device (dofs.py): no upload
        else:
            d["edofs"] = up(self.element_dofs)
        d["phi"], d["dphi"], d["W"], d["X"] = up(self._phi), up(self._dphi), up(self.W), up(self.X)
        d["tind"] = None if self.tind is None else up(self.tind.astype(np.int32))
        if not self._affine:
            d["mphi"], d["mdphi"] = up(self._mphi), up(self._mdphi)
        sp = _lib.SkbSpace()
        sp.dim = self.mesh.dim()
        sp.nnodes = self.mesh.t.shape[0]
        sp.mapping = _lib.SKB_MAP_AFFINE if self._affine else _lib.SKB_MAP_ISO_HEX1
        sp.nbs = self.nbs
        sp.ncomp = self.ncomp
        sp.nqp = self.nqp
        sp.npts = self.mesh.p.shape[1]
        sp.nel_total = self.mesh.nelements
        sp.p = p.data_ptr()
        sp.t = t.data_ptr()
        sp.tind = None if d["tind"] is None else d["tind"].data_ptr()
        sp.nel = self.nelems
        sp.phi, sp.dphi, sp.W, sp.X = (d["phi"].data_ptr(), d["dphi"].data_ptr(),
                                       d["W"].data_ptr(), d["X"].data_ptr())
        sp.mdphi = None if self._affine else d["mdphi"].data_ptr()
        sp.mphi = None if self._affine else d["mphi"].data_ptr()
        d["space"] = sp
        self._devcache[key] = d
        return d

    def update_points(self, p, host=True, adopt=False):
        """Move the mesh: new vertex coordinates ``p`` (``(dim, npts)`` numpy array or device
        tensor), same connectivity.  The reference has no such call - a moved mesh is a new
        ``Mesh`` + ``Basis`` (e.g. inside the Newton / time loops of docs/examples/ex10.py-style
        scripts) and every assembly re-derives DOFs, tables and the sparsity pattern.  Here
        the device copy of ``p`` is overwritten in place (all plans keep their pointers, a
        captured CUDA graph stays valid), cached geometry fields are dropped, and sparsity
        plans survive only where the next assembly can *validate* them: the fused P1 path
        compares the zero mask of every local matrix with the plan's and re-plans when the
        value-dependent pattern (coo_data.py:35) may have changed; all other cached plans are
        discarded, so those forms take the cold path once.  ``host=False`` skips refreshing
        ``mesh.p`` (the host copy then lags behind until the next call with ``host=True``).
        ``adopt=True`` (device tensors only): no copy at all - the basis and its mesh reference
        ``p`` itself from now on (the caller must keep it alive and unchanged until the next
        ``update_points``); this is how a loop alternating between coordinate buffers, or one
        CUDA graph per buffer, avoids the device-to-device copy."""
        torch = _torch()
        d = self._dev()
        src = p if torch.is_tensor(p) else torch.from_numpy(
            np.ascontiguousarray(p, dtype=np.float64))
        if tuple(src.shape) != tuple(d["p"].shape):
            raise ValueError("update_points: expected an array of shape {}".format(
                tuple(d["p"].shape)))
        if adopt:
            if not (torch.is_tensor(p) and p.device == d["p"].device and p.is_contiguous()
                    and p.dtype == torch.float64):
                raise ValueError("update_points(adopt=True) needs a contiguous float64 tensor "
                                 "on the basis' device")
            d["p"] = p
            d["space"].p = p.data_ptr()
            self.mesh._dev[str(d["device"])] = (p, d["t"])
        else:
            d["p"].copy_(src, non_blocking=True)
        if host:
            self.mesh.p[...] = src.cpu().numpy() if torch.is_tensor(p) else p
        self._fields.clear()
        if hasattr(self, "_doflocs"):
            del self._doflocs
        from . import fused2
        # value-independent (LinearForm scatter) and mask-validated plans stay valid
        keep = {k: v for k, v in self._plans.items() if k in ("linear", "by-mask")}
        for k, fp in self._plans.items():
            if isinstance(k, tuple) and k and k[0] == "fused" and getattr(fp, "version", 1) == 2:
                if fp.mode != 4:            # (4: the mass form has one arithmetic variant)
                    fp.mode = fused2.arithmetic_mode(d["p"], fp.w, fp.nqp)
                fp.p = d["p"]
                fp.unchecked = True
                keep[k] = fp
                keep[k[1]] = self._plans[k[1]]
        self._plans = keep

    @staticmethod
    def _stream():
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)

    def _tabulate(self, b=None, want=("grad",)):
        """Run skb_tabulate; returns dict of torch tensors."""
        torch = _torch()
        d = self._dev()
        dev, nel, nqp, dim = d["device"], self.nelems, self.nqp, self.mesh.dim()
        out = {}
        if "grad" in want:
            out["grad"] = torch.empty((dim, nel, nqp), dtype=torch.float64, device=dev)
        if "dx" in want:
            out["dx"] = torch.empty((nel, nqp), dtype=torch.float64, device=dev)
        if "x" in want:
            out["x"] = torch.empty((dim, nel, nqp), dtype=torch.float64, device=dev)
        if "detabs" in want:
            out["detabs"] = torch.empty((nel, nqp), dtype=torch.float64, device=dev)

        def ptr(k):
            return out[k].data_ptr() if k in out else None
        code = _lib.lib().skb_tabulate(C.byref(d["space"]), 0 if b is None else int(b),
                                       ptr("grad"), ptr("dx"), ptr("x"), ptr("detabs"),
                                       self._stream())
        _lib.check(code, "skb_tabulate")
        return out

    # -- materialised views (reference shapes), produced on demand ------------------------
    def _dx_dev(self):
        if "dx" not in self._fields:
            self._fields["dx"] = self._tabulate(want=("dx",))["dx"]
        return self._fields["dx"]

    @property
    def dx(self):
        """(nel, nqp) |detDF| * W, host numpy like the reference attribute."""
        return self._dx_dev().cpu().numpy()

    def _scalar_grad_dev(self, b):
        key = ("grad", b)
        if key not in self._fields:
            self._fields[key] = self._tabulate(b=b, want=("grad",))["grad"]
        return self._fields[key]

    def _basis_field_dev(self, i):
        """Device DiscreteField of local basis function ``i`` (value + grad),
        shaped like ElementH1.gbasis / ElementVector.gbasis output."""
        from .field import DiscreteField
        torch = _torch()
        d = self._dev()
        nel, nqp, dim, nc = self.nelems, self.nqp, self.mesh.dim(), self.ncomp
        b, n = divmod(i, nc)
        g = self._scalar_grad_dev(b)
        val = d["phi"][b].expand(nel, nqp)
        if nc == 1:
            # numpy: value is a stride-0 broadcast of phi (element_h1.py:15)
            return DiscreteField(val, g, lay="B")
        vv = torch.zeros((dim, nel, nqp), dtype=torch.float64, device=d["device"])
        gg = torch.zeros((dim, dim, nel, nqp), dtype=torch.float64, device=d["device"])
        vv[n] = val
        gg[n] = g
        return DiscreteField(vv, gg)

    @property
    def basis(self):
        """List of 1-tuples of (device) DiscreteFields, one per local basis
        function - the reference's ``basis.basis`` (cell_basis.py:101-102)."""
        return [(self._basis_field_dev(i),) for i in range(self.Nbfun)]

    def global_coordinates(self):
        from .field import DiscreteField
        if "x" not in self._fields:
            self._fields["x"] = self._tabulate(want=("x",))["x"]
        return DiscreteField(self._fields["x"])

    def mesh_parameters(self):
        from .field import DiscreteField
        if "h" not in self._fields:
            det = self._tabulate(want=("detabs",))["detabs"]
            self._fields["h"] = det ** (1. / self.mesh.dim())
        return DiscreteField(self._fields["h"], lay=self._dx_layout)

    @property
    def _dx_layout(self):
        """numpy layout of the reference's ``dx``: MappingAffine.detDF is
        ``np.tile(detA, (nqp, 1)).T`` (mapping_affine.py:205-211), i.e.
        Fortran-ordered; the isoparametric one is C-ordered."""
        return "F" if self._affine else "C"

    def default_parameters(self):
        """``w.x`` and ``w.h`` (cell_basis.py:124-141) as device fields."""
        return {"x": self.global_coordinates(), "h": self.mesh_parameters()}

    def interpolate(self, w):
        """Field of the FE function with DOF vector ``w`` at the quadrature
        points: sum_i w[element_dofs[i]] * basis_i, value and gradient
        (abstract_basis.py:271-322), as a device DiscreteField."""
        from .field import DiscreteField
        torch = _torch()
        d = self._dev()
        if not torch.is_tensor(w):
            w = torch.as_tensor(np.asarray(w, dtype=np.float64), device=d["device"])
        if w.shape[0] != self.N:
            raise ValueError("Input array has wrong size.")
        val = grd = None
        ed = d["edofs"].long()
        for i in range(self.Nbfun):
            f = self._basis_field_dev(i)
            coef = w[ed[i]][:, None]
            tv = coef * f.t
            tg = coef * f.grad.t
            val = tv if val is None else val + tv
            grd = tg if grd is None else grd + tg
        return DiscreteField(val, grd)

    def __repr__(self):
        return ("<skfem_b200 {}({}, {}) object>\n  Number of elements: {}\n"
                "  Number of DOFs: {}\n  Size: {} B").format(
            type(self).__name__, type(self.mesh).__name__, type(self.elem).__name__,
            self.nelems, self.N, 0)


Basis = CellBasis
