"""Meshes: ``p`` (dim, nverts) float64 and ``t`` (nnodes, nel) int32 plus the
lazily built topology needed for DOF numbering.

Host side (numpy) by design: mesh generation is not the path being accelerated
(SURVEY.md section 2 row 14).  Layout, dtypes and the vertex/element ordering of
the constructors follow the reference so that ``p``/``t`` are bit-identical:

* container + dtype normalisation        skfem/mesh/mesh.py:27-28,544-607
* edges/facets = unique sorted tuples    skfem/mesh/mesh.py:1065-1082
* MeshTet.init_tensor (6 Kuhn tets/cell) skfem/mesh/mesh_tet_1.py:326-393
* MeshHex.init_tensor                    skfem/mesh/mesh_hex_1.py:97-155
* MeshTri defaults / uniform refinement  skfem/mesh/mesh_tri_1.py:14-28,209-227
"""
from __future__ import annotations

import numpy as np

from .element import ElementTriP1, ElementTetP1, ElementHex1


def _cuda_ready():
    try:
        import torch
        return torch.cuda.is_available()
    except ImportError:
        return False


def _pair_plan(tv, tu, adj, vmax, nrows):
    """Sorted unique (row, col) pairs - rows from ``tv`` (nbv, nel), cols from ``tu`` (nbu,
    nel), kept where ``tu[j] > vmax[i]`` for j in ``adj[i]`` - by the row-bucket plan builder
    (csrc/skb_plan_rows.cu, skb_entity_masks): (row_of_slot, col_of_slot, slot_of_entry) or
    None when a limit of that path is hit."""
    import ctypes as C
    import torch
    from . import _lib
    from .form import build_plan
    lib = _lib.lib()
    nbv, nel = int(tv.shape[0]), int(tv.shape[1])
    nbu = int(tu.shape[0])
    dev = tv.device
    mask = torch.empty(nbv * nel, dtype=torch.int32, device=dev)
    adj_c = (C.c_uint32 * nbv)(*[int(a) for a in adj])
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.skb_entity_masks(tu.data_ptr(), nbu, nbv, nel, vmax.data_ptr(), adj_c,
                                    mask.data_ptr(), stream), "skb_entity_masks")
    plan = build_plan(tv, tu, nel, (nrows, nrows), None, mask=mask)
    if plan is None:
        return None
    slot = torch.full((nbu * nbv * nel,), -1, dtype=torch.int32, device=dev)
    _lib.check(lib.skb_plan_slot_of_entry(plan.segptr.data_ptr(), plan.perm.data_ptr(), plan.nnz,
                                          slot.data_ptr(), stream), "skb_plan_slot_of_entry")
    counts = (plan.indptr[1:] - plan.indptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(nrows, device=dev, dtype=torch.int64), counts)
    return rows, plan.indices.long(), slot


def _edges_rows(tt, indices, nv):
    """Edges (k = 2) of a mesh by the library's own kernels: entities (2, nedges) sorted
    lexicographically and the incidence (len(indices), nel), both device int64 tensors."""
    import torch
    nn, nel = int(tt.shape[0]), int(tt.shape[1])
    adj = [0] * nn
    for a, b in indices:
        adj[a] |= 1 << b
        adj[b] |= 1 << a
    res = _pair_plan(tt, tt, adj, tt, nv)
    if res is None:
        return None
    rows, cols, slot = res
    e = torch.arange(nel, device=tt.device, dtype=torch.int64)
    inc = []
    for a, b in indices:      # entry (j * nn + i) * nel + e with i = the smaller vertex
        k = torch.where(tt[a] < tt[b], b * nn + a, a * nn + b).long() * nel + e
        inc.append(slot[k].long())
    return torch.stack([rows, cols]), torch.stack(inc)


def _entities_to_host(ent32, inc32, tdt):
    """Device entity / incidence tensors -> the host arrays of ``Mesh.build_entities``."""
    return (np.ascontiguousarray(ent32.cpu().numpy().astype(tdt, copy=False)),
            inc32.cpu().numpy().astype(np.int64))


def _build_entities_device(t, indices, nv, sort, keep=None):
    """``Mesh.build_entities`` on the GPU; same results as the host path,
    returned as host arrays because the numbering API is host numpy.  Edges of any mesh and the
    triangular facets of tetrahedra go through the library's own row-bucket sort (the unique
    sorted pairs are a CSR pattern, csrc/skb_plan_rows.cu); quadrilateral facets and
    ``sort=False`` numberings use torch's sort / unique on packed integer keys."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    k, n = len(indices[0]), t.shape[1]
    tdt = t.dtype
    tt32 = torch.from_numpy(np.ascontiguousarray(t, dtype=np.int32)).to(dev)

    def host(ent, inc):
        inc32 = inc.to(torch.int32).contiguous()
        ent32 = ent.to(torch.int32).contiguous()
        if keep is not None:
            keep["inc"], keep["ent"] = inc32, ent32    # device copies (Dofs: no re-upload)
            if keep.get("defer"):                      # host arrays on demand (_entities_to_host)
                return None, None
        return _entities_to_host(ent32, inc32, tdt)
    if sort and k == 2:
        res = _edges_rows(tt32, [tuple(ix) for ix in indices], nv)
        if res is not None:
            return host(*res)
    if sort and k == 3 and t.shape[0] == 4:
        # facet (a < b < c) = (edge of the two smallest vertices, largest vertex)
        ledges = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
        er = _edges_rows(tt32, ledges, nv)
        if er is not None:
            edges, t2e = er
            nedges = int(edges.shape[1])
            adj = [sum(1 << j for j in range(4) if j not in le) for le in ledges]
            vmax = torch.stack([torch.maximum(tt32[a], tt32[b]) for a, b in ledges]).contiguous()
            res = _pair_plan(t2e.to(torch.int32).contiguous(), tt32, adj, vmax, nedges)
            if res is not None:
                rows, cols, slot = res
                ent = torch.stack([edges[0][rows], edges[1][rows], cols])
                eidx = {le: i for i, le in enumerate(ledges)}
                e = torch.arange(n, device=dev, dtype=torch.int64)
                inc = []
                for ix in indices:
                    ix = tuple(int(v) for v in ix)
                    vals = torch.stack([tt32[v] for v in ix])            # (3, nel)
                    order = torch.argsort(vals, dim=0)                    # local positions
                    loc = torch.tensor(ix, device=dev)[order]             # (3, nel) local ids
                    lo, mid, hi = loc[0], loc[1], loc[2]
                    le = torch.zeros(n, dtype=torch.int64, device=dev)
                    for (a, b), i in eidx.items():
                        le = torch.where(((lo == a) & (mid == b)) | ((lo == b) & (mid == a)),
                                         torch.full_like(le, i), le)
                    inc.append(slot[(hi * 6 + le) * n + e].long())
                return host(ent, torch.stack(inc))
    tt = tt32.long()
    stacked = torch.cat([tt[list(ix)] for ix in indices], dim=1)          # (k, n * len(indices))
    canon = torch.sort(stacked, dim=0).values
    if k == 4 and float(nv) ** 4 >= 2.0 ** 62:
        # two levels: dense rank of the leading pair, then (rank, trailing pair) - both fit
        hi_key = canon[0] * nv + canon[1]
        uhi, rhi = torch.unique(hi_key, sorted=True, return_inverse=True)
        key = (rhi * nv + canon[2]) * nv + canon[3]
        ukey, inverse = torch.unique(key, sorted=True, return_inverse=True)
        incidence = inverse.reshape(len(indices), n)
        if not sort:
            first = torch.full((ukey.shape[0],), inverse.shape[0], dtype=torch.int64, device=dev)
            first.scatter_reduce_(0, inverse, torch.arange(inverse.shape[0], device=dev), "amin")
            return host(stacked[:, first], incidence)
        ent = torch.empty((4, ukey.shape[0]), dtype=torch.int64, device=dev)
        ent[3] = ukey % nv
        rest = ukey // nv
        ent[2] = rest % nv
        lead = uhi[rest // nv]
        ent[1] = lead % nv
        ent[0] = lead // nv
        return host(ent, incidence)
    if float(nv) ** k >= 2.0 ** 62:
        return None                                   # keys do not pack: host path
    key = canon[0]
    for r in range(1, k):
        key = key * nv + canon[r]
    del canon
    ukey, inverse = torch.unique(key, sorted=True, return_inverse=True)
    del key
    incidence = inverse.reshape(len(indices), n)
    if not sort:   # representative = first occurrence, like np.unique(return_index=True)
        first = torch.full((ukey.shape[0],), inverse.shape[0], dtype=torch.int64, device=dev)
        first.scatter_reduce_(0, inverse, torch.arange(inverse.shape[0], device=dev), "amin")
        return host(stacked[:, first], incidence)
    ent = torch.empty((k, ukey.shape[0]), dtype=torch.int64, device=dev)
    for r in range(k - 1, -1, -1):
        ent[r] = ukey % nv
        ukey = ukey // nv
    return host(ent, incidence)


class OrientedBoundary(np.ndarray):
    """Facet indices plus one orientation flag per facet (the role of
    skfem/generic_utils.py:16-28): ``ori[k]`` selects the row of ``f2t[:, find[k]]`` - the
    element - that traces and outward normals are taken from (FacetBasis,
    facet_basis.py:84-89).  Behaves like the plain index array everywhere else."""
    ori = None

    def __new__(cls, indices, ori):
        self = np.asarray(indices).view(cls)
        flags = np.array(ori, dtype=int).reshape(-1)
        if flags.shape[0] != self.shape[0]:
            raise ValueError("OrientedBoundary: one orientation per facet")
        self.ori = flags
        return self

    def __array_finalize__(self, source):
        # views and copies keep the flags of the array they come from
        if source is not None and self.ori is None:
            self.ori = getattr(source, "ori", None)


class Mesh:
    elem = None          # geometry element (class)
    affine = False
    sort_t = False

    def __init__(self, doflocs=None, t=None, validate=True, **_ignored):
        if doflocs is None:
            doflocs, t = self._default()
        t = np.asarray(t)
        if self.sort_t:
            t = np.sort(t, axis=0)
        self.doflocs = np.ascontiguousarray(np.asarray(doflocs, dtype=np.float64))
        self.t = np.ascontiguousarray(t.astype(np.int32, copy=False))
        if self.t.shape[0] != self.refdom.nnodes:
            raise ValueError("t must have {} rows".format(self.refdom.nnodes))
        if self.doflocs.shape[0] != self.refdom.dim():
            raise ValueError("p must have {} rows".format(self.refdom.dim()))
        self._dev = {}

    # -- basic queries -------------------------------------------------------
    @property
    def p(self):
        return self.doflocs

    @property
    def refdom(self):
        return self.elem.refdom

    def dim(self):
        return self.refdom.dim()

    @property
    def nelements(self):
        return self.t.shape[1]

    @property
    def nvertices(self):
        """``max(t) + 1`` like the reference (mesh/mesh.py:71-73).  For large
        meshes the reduction runs on the GPU over the connectivity that has to
        be uploaded anyway (a 24 M-entry host max costs ~6 ms)."""
        if not hasattr(self, "_nvertices"):
            nv = None
            if self.t.size > (1 << 20):
                try:
                    import torch
                    if torch.cuda.is_available():
                        dev = torch.device("cuda", torch.cuda.current_device())
                        nv = int(self.device_arrays(dev)[1].max()) + 1
                except ImportError:
                    pass
            self._nvertices = int(np.max(self.t)) + 1 if nv is None else nv
        return self._nvertices

    @property
    def nnodes(self):
        return self.t.shape[0]

    def _nentities(self, what):
        names = {"facets": "_facets", "edges": "_edges"}[what]
        if not hasattr(self, names) and not hasattr(self, names + "_dev"):
            self._init_entities(what, getattr(self.refdom, what),
                                self._sort_facets if what == "facets" else True)
        if hasattr(self, names):
            return getattr(self, names).shape[1]
        return int(getattr(self, names + "_dev").shape[1])

    @property
    def nfacets(self):
        return self._nentities("facets")

    @property
    def nedges(self):
        if self.refdom.edges is None:
            raise NotImplementedError
        return self._nentities("edges")

    def __repr__(self):
        return "<skfem_b200 {} object>\n  Number of elements: {}\n  Number of vertices: {}".format(
            type(self).__name__, self.nelements, self.nvertices)

    # -- topology --------------------------------------------------------------
    @staticmethod
    def build_entities(t, indices, sort=True, _keep=None):
        """Lower-dimensional entities as the lexicographically sorted unique
        columns of the per-element sorted vertex tuples, and the element ->
        entity incidence.  ``_keep`` (dict): receives the device copy of the incidence
        when the numbering ran on the GPU."""
        k = len(indices[0])
        nv = int(t.max()) + 1 if t.size else 1
        if t.shape[1] >= (1 << 16) and _cuda_ready():
            res = _build_entities_device(t, indices, nv, sort, _keep)
            if res is not None:
                return res
        if nv ** k < (1 << 62):
            # a sorted vertex tuple packs into one int64 whose order is the
            # lexicographic order of the tuple: 1-D unique instead of numpy's
            # structured-dtype argsort (2.5x faster on the host); with a CUDA
            # device and a large mesh the numbering runs there (above)
            stacked = np.hstack([t[ix] for ix in indices])
            canon = np.sort(stacked, axis=0).astype(np.int64)
            key = canon[0]
            for r in range(1, k):
                key = key * nv + canon[r]
            ukey, first, inverse = np.unique(key, return_index=True, return_inverse=True)
            incidence = inverse.reshape((len(indices), t.shape[1]))
            if not sort:
                return np.ascontiguousarray(stacked[:, first]), incidence
            ent = np.empty((k, ukey.shape[0]), dtype=t.dtype)
            for r in range(k - 1, -1, -1):
                ent[r] = ukey % nv
                ukey = ukey // nv
            return ent, incidence
        stacked = np.hstack([t[ix] for ix in indices])
        canon = np.sort(stacked, axis=0)
        canon, first, inverse = np.unique(canon, axis=1, return_index=True,
                                          return_inverse=True)
        incidence = inverse.reshape((len(indices), t.shape[1]))
        if sort:
            return np.ascontiguousarray(canon), incidence
        return np.ascontiguousarray(stacked[:, first]), incidence

    _sort_facets = True

    # Entities numbered on the GPU stay there (``_facets_dev`` / ``_t2f_dev`` ...: Dofs builds
    # element_dofs from them without a round trip); the host arrays of the public attributes
    # are fetched on first access.
    def _init_entities(self, what, indices, sort):
        keep = {"defer": True}
        ent, inc = self.build_entities(self.t, indices, sort=sort, _keep=keep)
        names = {"facets": ("_facets", "_t2f"), "edges": ("_edges", "_t2e")}[what]
        if ent is None:                               # deferred: device tensors only
            setattr(self, names[0] + "_dev", keep["ent"])
            setattr(self, names[1] + "_dev", keep["inc"])
        else:
            setattr(self, names[0], ent)
            setattr(self, names[1], inc)
            setattr(self, names[1] + "_dev", keep.get("inc"))

    def _host_entities(self, what):
        names = {"facets": ("_facets", "_t2f"), "edges": ("_edges", "_t2e")}[what]
        if not hasattr(self, names[0]):
            if not hasattr(self, names[0] + "_dev"):
                self._init_entities(what, getattr(self.refdom, what),
                                    self._sort_facets if what == "facets" else True)
            if not hasattr(self, names[0]):
                ent, inc = _entities_to_host(getattr(self, names[0] + "_dev"),
                                             getattr(self, names[1] + "_dev"), self.t.dtype)
                setattr(self, names[0], ent)
                setattr(self, names[1], inc)

    def _init_facets(self):
        self._host_entities("facets")

    def _init_edges(self):
        self._host_entities("edges")

    @property
    def facets(self):
        if not hasattr(self, "_facets"):
            self._init_facets()
        return self._facets

    @property
    def t2f(self):
        if not hasattr(self, "_t2f"):
            self._init_facets()
        return self._t2f

    @property
    def edges(self):
        if not hasattr(self, "_edges"):
            self._init_edges()
        return self._edges

    @property
    def t2e(self):
        if not hasattr(self, "_t2e"):
            self._init_edges()
        return self._t2e

    @property
    def f2t(self):
        """(2, nfacets): the one or two elements sharing each facet (-1)."""
        if not hasattr(self, "_f2t"):
            nf = self.nfacets
            flat = self.t2f.flatten(order='C')
            owner = np.tile(np.arange(self.nelements), self.t2f.shape[0])
            out = np.full((2, nf), -1, dtype=np.int32)
            order = np.argsort(flat, kind='stable')
            fs, es = flat[order], owner[order]
            start = np.r_[True, fs[1:] != fs[:-1]]
            last = np.r_[fs[1:] != fs[:-1], True]
            out[0, fs[start]] = es[start]
            two = last & ~start
            out[1, fs[two]] = es[two]
            self._f2t = out
        return self._f2t

    def boundary_facets(self):
        return np.nonzero(self.f2t[1] == -1)[0].astype(np.int32)

    def boundary_nodes(self):
        return np.unique(self.facets[:, self.boundary_facets()])

    def facet_edges(self, ix):
        """Edges (3-D) lying on the facets ``ix`` - the reference's
        ``np.unique(f2e[:, ix])`` (mesh.py:495-511)."""
        bf = self.facets[:, np.asarray(ix, dtype=np.int64)]
        n = bf.shape[0]
        # facet vertex loops: every consecutive pair is an edge (triangles: all
        # three pairs; quads of a hex: the four sides)
        pairs = np.hstack([np.sort(bf[[a, (a + 1) % n]], axis=0) for a in range(n)])
        nv = self.nvertices
        key = self.edges[0].astype(np.int64) * nv + self.edges[1]
        ckey = np.unique(pairs[0].astype(np.int64) * nv + pairs[1])
        return np.nonzero(np.isin(key, ckey))[0].astype(np.int32)

    def boundary_edges(self):
        """Edges (3-D) lying on boundary facets."""
        return self.facet_edges(self.boundary_facets())

    boundaries = None  # optional {name: facet indices}

    def facets_satisfying(self, test, boundaries_only=False, normal=None):
        """Facets whose midpoints satisfy ``test`` (mesh.py:426-456); with ``normal`` an
        ``OrientedBoundary``: orientation 1 where ``normal`` points against the outward
        normal of the facet's first element."""
        midp = self.p[:, self.facets].mean(axis=1)
        facets = np.nonzero(test(midp))[0].astype(np.int32)
        if boundaries_only:
            facets = np.intersect1d(facets, self.boundary_facets())
        if normal is not None:
            ori = 1 * (np.dot(np.asarray(normal, dtype=np.float64),
                              self._outward_normals(facets)) < 0)
            return OrientedBoundary(facets, ori)
        return facets

    def _outward_normals(self, facets):
        """(dim, len(facets)) outward normal directions (not normalised) of the facets with
        respect to their first element ``f2t[0]`` - the direction of
        ``MappingAffine.normals`` (mapping_affine.py:248-281), from the vertices alone."""
        if not self.affine:
            raise NotImplementedError("oriented facet sets: affine meshes only")
        fv = self.p[:, self.facets[:, facets]]                  # (dim, nodes, nf)
        if self.dim() == 2:
            e = fv[:, 1] - fv[:, 0]
            n = np.array([e[1], -e[0]])
        else:
            a, b = fv[:, 1] - fv[:, 0], fv[:, 2] - fv[:, 0]
            n = np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2],
                          a[0] * b[1] - a[1] * b[0]])
        centre = self.p[:, self.t[:, self.f2t[0, facets]]].mean(axis=1)
        inward = np.sum(n * (centre - fv[:, 0]), axis=0) > 0
        return np.where(inward, -n, n)

    def facets_around(self, elements, flip=False):
        """The oriented facets around a set of elements (mesh.py:458-479): traces and outward
        normals from inside the set, or from outside it with ``flip=True``."""
        elements = self.normalize_elements(elements)
        facets, counts = np.unique(self.t2f[:, elements], return_counts=True)
        facets = facets[counts == 1]
        member = np.isin(self.f2t[:, facets], elements).T
        ori = np.nonzero(~member if flip else member)[1].astype(np.int32)
        return OrientedBoundary(facets, ori)

    def with_boundaries(self, boundaries, boundaries_only=True):
        """Copy of the mesh with named boundaries: ``{name: facet indices | test on
        facet midpoints}`` (mesh.py:791-830)."""
        out = type(self)(self.doflocs, self.t)
        named = dict(self.boundaries or {})
        for name, spec in boundaries.items():
            named[name] = (self.facets_satisfying(spec, boundaries_only=boundaries_only)
                           if callable(spec) else np.asarray(spec, dtype=np.int32))
        out.boundaries = named
        out.subdomains = self.subdomains
        for attr in ("_facets", "_t2f", "_edges", "_t2e", "_f2t", "_nvertices", "_facets_dev",
                     "_t2f_dev", "_edges_dev", "_t2e_dev"):
            if hasattr(self, attr):
                setattr(out, attr, getattr(self, attr))
        return out

    def normalize_facets(self, facets):
        """Array of facet indices from an index, an array, a list of criteria,
        a callable on facet midpoints or a boundary name (mesh.py:1291-1329)."""
        if isinstance(facets, (int, np.integer)):
            return np.array([facets])
        if isinstance(facets, np.ndarray):
            return facets
        if facets is None:
            return self.boundary_facets()
        if isinstance(facets, (tuple, list, set)):
            return np.unique(np.concatenate([self.normalize_facets(f) for f in facets]))
        if callable(facets):
            return self.facets_satisfying(facets)
        if isinstance(facets, str):
            if self.boundaries is not None and facets in self.boundaries:
                return self.boundaries[facets]
            raise ValueError("Boundary '{}' not found.".format(facets))
        raise NotImplementedError

    def normalize_elements(self, elements):
        """Element indices from an index, an index / boolean array, a test on element
        midpoints, a subdomain name, or a list / set of those (mesh.py:1331-1368)."""
        if isinstance(elements, bool):
            if elements:
                return np.arange(self.nelements, dtype=np.int32)
            raise NotImplementedError
        if isinstance(elements, (int, np.integer)):
            return np.array([elements], dtype=np.int32)
        if callable(elements):
            return self.elements_satisfying(elements)
        if isinstance(elements, str):
            if self.subdomains is not None and elements in self.subdomains:
                return self.subdomains[elements]
            raise ValueError("Subdomain '{}' not found.".format(elements))
        if isinstance(elements, (tuple, list, set)):
            if len(elements) == 0:
                return np.zeros(0, dtype=np.int32)
            return np.unique(np.concatenate([self.normalize_elements(e) for e in elements]))
        arr = np.asarray(elements)
        if arr.dtype == bool:
            arr = np.nonzero(arr)[0]
        return arr.astype(np.int32)

    # -- selections on nodes / elements (mesh.py:402-424,476-493,251-275) ---------
    subdomains = None  # optional {name: element indices}

    def interior_nodes(self):
        return np.setdiff1d(np.arange(self.p.shape[1]), self.boundary_nodes())

    def nodes_satisfying(self, test, boundaries_only=False):
        nodes = np.nonzero(test(self.p))[0].astype(np.int32)
        return np.intersect1d(nodes, self.boundary_nodes()) if boundaries_only else nodes

    def normalize_nodes(self, nodes):
        """Vertex indices from an index, an array, a test on vertex coordinates, a point given
        as a tuple of coordinates, or a list / set of those (mesh.py:1259-1289)."""
        if isinstance(nodes, tuple):
            pt = np.array(list(nodes))[:, None]
            return self.nodes_satisfying(lambda x: np.linalg.norm(x - pt, axis=0) < 1e-12)
        if isinstance(nodes, np.ndarray):
            return nodes
        if isinstance(nodes, (list, set)):
            return np.unique(np.concatenate([self.normalize_nodes(n) for n in nodes]))
        if callable(nodes):
            return self.nodes_satisfying(nodes)
        raise NotImplementedError    # bare integers: like the reference, pass an array

    def elements_satisfying(self, test):
        """Elements whose midpoint (mean of the vertices) satisfies ``test``."""
        return np.nonzero(test(self.p[:, self.t].mean(axis=1)))[0].astype(np.int32)

    def with_subdomains(self, subdomains):
        """Copy of the mesh with named element sets ``{name: indices | test on midpoints}``."""
        out = self.with_boundaries({})
        out.boundaries = self.boundaries
        named = dict(self.subdomains or {})
        for name, spec in subdomains.items():
            named[name] = self.elements_satisfying(spec) if callable(spec) else spec
        out.subdomains = named
        return out

    def params(self):
        """Per element, the length of its longest edge (mesh_2d.py:12-17, mesh_3d.py:13-17)."""
        ents, t2x = (self.edges, self.t2e) if self.dim() == 3 else (self.facets, self.t2f)
        length = np.linalg.norm(np.diff(self.p[:, ents], axis=1), axis=0)[0]
        return length[t2x].max(axis=0)

    def param(self):
        """Mesh parameter h: the longest edge of the mesh."""
        return np.max(self.params())

    # -- geometry --------------------------------------------------------------
    def _mapping(self):
        from .mapping import MappingAffine, MappingIsoparametric
        if not hasattr(self, "_cached_mapping"):
            self._cached_mapping = (MappingAffine(self) if self.affine
                                    else MappingIsoparametric(self, self.elem()))
        return self._cached_mapping

    def mapping(self):
        return self._mapping()

    def refined(self, times_or_ix=1):
        m = self
        for _ in range(int(times_or_ix)):
            m = m._uniform()
        return m

    def _uniform(self):
        raise NotImplementedError("uniform refinement is not implemented for "
                                  + type(self).__name__)

    def morphed(self, *args):
        """Apply one function per coordinate (``None`` keeps it)."""
        p = self.p.copy()
        for i, f in enumerate(args):
            if f is not None:
                p[i] = f(self.p)
        return type(self)(p, self.t)

    def translated(self, diffs):
        p = self.p.copy()
        for i, d in enumerate(diffs):
            p[i] += d
        return type(self)(p, self.t)

    def scaled(self, factors):
        p = self.p.copy()
        if np.isscalar(factors):
            factors = self.p.shape[0] * (factors,)
        for i, f in enumerate(factors):
            p[i] *= f
        return type(self)(p, self.t)

    # -- device mirror -----------------------------------------------------------
    def device_arrays(self, device):
        """(p, t) as torch tensors on ``device``, uploaded once per mesh."""
        import torch
        key = str(device)
        if key not in self._dev:
            # pinned host arrays copy asynchronously (stream ordered with the kernels)
            self._dev[key] = (
                torch.from_numpy(self.doflocs).to(device, non_blocking=True),
                torch.from_numpy(self.t).to(device, non_blocking=True),
            )
        return self._dev[key]


def _tensor_mesh_device(cls, x, y, z, corners):
    """``init_tensor`` on the GPU for large grids (csrc/skb_mesh.cu): the same p / t as the host
    generator, the host copies fetched into pinned memory and the device copies kept as the
    mesh's device arrays (no upload later).  None if no GPU / small grid / int32 overflow."""
    npx, npy, npz = len(x), len(y), len(z)
    ncells = (npx - 1) * (npy - 1) * (npz - 1)
    if ncells * len(corners) < (1 << 16) or not _cuda_ready():
        return None
    import ctypes as C
    import torch
    from . import _lib
    dev = torch.device("cuda", torch.cuda.current_device())
    xs, ys, zs = (torch.from_numpy(np.sort(np.asarray(v, dtype=np.float64))).to(dev)
                  for v in (x, y, z))
    nn, npts = len(corners[0]), npx * npy * npz
    p = torch.empty((3, npts), dtype=torch.float64, device=dev)
    t = torch.empty((nn, ncells * len(corners)), dtype=torch.int32, device=dev)
    tab = (C.c_int32 * (len(corners) * nn))(*[int(v) for row in corners for v in row])
    code = _lib.lib().skb_mesh_tensor(xs.data_ptr(), ys.data_ptr(), zs.data_ptr(), npx, npy, npz,
                                      len(corners), nn, tab, p.data_ptr(), t.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if code == _lib.SKB_ETOOBIG:
        return None
    _lib.check(code, "skb_mesh_tensor")
    ph = torch.empty(p.shape, dtype=p.dtype, pin_memory=True)
    th = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    ph.copy_(p, non_blocking=True)
    th.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    m = cls(ph.numpy(), th.numpy())
    m._pinned = (ph, th)                       # keeps the pinned buffers alive
    m._dev[str(dev)] = (p, t)
    m._nvertices = npts
    return m


def _tensor_grid(x, y, z):
    """Vertices of a tensor grid (vertex index = iy + npy*ix + npy*npx*iz) and
    the eight corner rows of each cell, cells enumerated the same way."""
    npx, npy, npz = len(x), len(y), len(z)
    X, Y, Z = np.meshgrid(np.sort(x), np.sort(y), np.sort(z))
    p = np.vstack((X.flatten('F'), Y.flatten('F'), Z.flatten('F')))
    ix = np.arange(npx * npy * npz, dtype=np.int64).reshape(npy, npx, npz, order='F')
    ne = (npx - 1) * (npy - 1) * (npz - 1)
    cell = ix[:-1, :-1, :-1].reshape(ne, order='F')
    sy, sx, sz = 1, npy, npy * npx
    offs = [0, sy, sx, sz, sy + sx, sy + sz, sx + sz, sy + sx + sz]
    corners = np.stack([cell + o for o in offs])
    return p, corners, ne


class MeshTri(Mesh):
    elem = ElementTriP1
    affine = True
    sort_t = True

    @staticmethod
    def _default():
        return (np.array([[0., 0.], [1., 0.], [0., 1.], [1., 1.]]).T,
                np.array([[0, 1, 2], [1, 3, 2]]).T)

    @classmethod
    def init_tensor(cls, x, y):
        npx, npy = len(x), len(y)
        X, Y = np.meshgrid(np.sort(x), np.sort(y))
        p = np.vstack((X.flatten('F'), Y.flatten('F')))
        ix = np.arange(npx * npy, dtype=np.int64).reshape(npy, npx, order='F')
        nt = (npx - 1) * (npy - 1)
        c = ix[:-1, :-1].reshape(nt, order='F')
        # two triangles per cell sharing the diagonal (corner 0 -> corner 3)
        t = np.hstack((np.stack((c, c + 1, c + npy + 1)),
                       np.stack((c, c + npy, c + npy + 1))))
        return cls(p, t)

    def _uniform(self):
        """Red refinement: new vertices at facet (edge) midpoints."""
        p, t, sz = self.p, self.t, self.p.shape[1]
        mid = self.t2f + sz
        newp = np.hstack((p, p[:, self.facets].mean(axis=1)))
        newt = np.hstack((
            np.vstack((t[0], mid[0], mid[2])),
            np.vstack((t[1], mid[0], mid[1])),
            np.vstack((t[2], mid[2], mid[1])),
            np.vstack((mid[0], mid[1], mid[2])),
        ))
        return type(self)(newp, newt)


class MeshTet(Mesh):
    elem = ElementTetP1
    affine = True

    @staticmethod
    def _default():
        p = np.array([[0., 0., 0.], [0., 0., 1.], [0., 1., 0.], [1., 0., 0.],
                      [0., 1., 1.], [1., 0., 1.], [1., 1., 0.], [1., 1., 1.]]).T
        t = np.array([[0, 1, 2, 3], [3, 5, 1, 7], [2, 3, 6, 7], [2, 3, 1, 7],
                      [1, 2, 4, 7]]).T
        return p, t

    @classmethod
    def init_tensor(cls, x, y, z):
        """Tensor-product grid, each cell split into the six Kuhn tetrahedra
        that share the body diagonal (corner 0 -> corner 7); the element order
        is type-major: all cells' first tet, then all second tets, ..."""
        kuhn = ([0, 1, 5, 7], [0, 1, 4, 7], [0, 2, 4, 7],
                [0, 3, 5, 7], [0, 2, 6, 7], [0, 3, 6, 7])
        m = _tensor_mesh_device(cls, x, y, z, kuhn)
        if m is not None:
            return m
        p, c, _ = _tensor_grid(x, y, z)
        return cls(p, np.hstack([c[rows] for rows in kuhn]))

    @classmethod
    def init_ball(cls, nrefs=3):
        """Unit ball (mesh_tet_1.py:396-428): the octahedron with vertices 0, +e_i (1..3),
        -e_i (4..6) cut into one tet per octant, refined ``nrefs`` times; after every
        refinement the boundary nodes are pushed out onto the unit sphere."""
        p = np.vstack((np.zeros((1, 3)), np.eye(3), 0. - np.eye(3))).T   # 0. - x: no -0.0
        octants = [[0, 1, 2, 3], [0, 4, 5, 6], [0, 1, 2, 6], [0, 1, 3, 5],
                   [0, 2, 3, 4], [0, 4, 5, 3], [0, 4, 6, 2], [0, 5, 6, 1]]
        m = cls(p, np.array(octants, dtype=np.int32).T)
        for _ in range(nrefs):
            m = m.refined()
            q = m.p.copy()
            shell = m.boundary_nodes()
            q[:, shell] = q[:, shell] / np.linalg.norm(q[:, shell], axis=0)
            m = cls(q, m.t)
        return m

    # local edges (rows of t2e) that meet at local vertex 0..3
    _CORNER_EDGES = ((0, 2, 3), (0, 1, 4), (1, 2, 5), (3, 4, 5))
    # the inner octahedron has three diagonals, each joining the midpoints of two opposite
    # edges; cut along diagonal (a, b) it falls into four tets (a, b, x, y) with (x, y) from
    # the ring of the remaining four midpoints
    _OCTA = (((2, 4), ((0, 1), (0, 3), (1, 5), (3, 5))),
             ((1, 3), ((0, 4), (4, 5), (5, 2), (2, 0))),
             ((0, 5), ((1, 4), (4, 3), (3, 2), (2, 1))))

    def _uniform(self):
        """Regular 1:8 refinement (the scheme of mesh_tet_1.py:84-126): vertices at the edge
        midpoints, one child per corner, the inner octahedron cut along its shortest diagonal
        - measured, like the reference, in the x-y plane only; ties go to the later diagonal.
        Children are ordered corner-major, then ring-position-major / diagonal-minor."""
        p, t = self.p, self.t
        mid = self.t2e + p.shape[1]                       # vertex id of every local edge midpoint
        newp = np.hstack((p, p[:, self.edges].mean(axis=1)))

        def d2(a, b):                                     # squared x-y distance of two midpoints
            return ((newp[0, mid[a]] - newp[0, mid[b]]) ** 2
                    + (newp[1, mid[a]] - newp[1, mid[b]]) ** 2)
        d = [d2(a, b) for (a, b), _ in self._OCTA]
        pick = (np.logical_and(d[0] < d[1], d[0] < d[2]),
                np.logical_and(~(d[0] < d[1]), d[1] < d[2]),
                np.logical_and(~(d[0] < d[2]), ~(d[1] < d[2])))
        children = [np.vstack((t[v],) + tuple(mid[e] for e in self._CORNER_EDGES[v]))
                    for v in range(4)]
        for ring in range(4):
            for ((a, b), pairs), sel in zip(self._OCTA, pick):
                x, y = pairs[ring]
                children.append(np.vstack((mid[a, sel], mid[b, sel], mid[x, sel], mid[y, sel])))
        return type(self)(newp, np.hstack(children))


class MeshHex(Mesh):
    elem = ElementHex1
    affine = False
    _sort_facets = False    # hex facets keep their loop order (mesh_hex_1.py:49-55)

    @staticmethod
    def _default():
        p = np.array([[0., 0., 0.], [0., 0., 1.], [0., 1., 0.], [1., 0., 0.],
                      [0., 1., 1.], [1., 0., 1.], [1., 1., 0.], [1., 1., 1.]]).T
        return p, np.arange(8).reshape(8, 1)

    @classmethod
    def init_tensor(cls, x, y, z):
        m = _tensor_mesh_device(cls, x, y, z, (list(range(8)),))
        if m is not None:
            return m
        p, c, _ = _tensor_grid(x, y, z)
        return cls(p, c)

    def _uniform(self):
        """1:8 refinement (the scheme of mesh_hex_1.py:57-95): new nodes at the edge midpoints,
        facet centres and cell centres, numbered in that order after the old vertices.  Child
        k keeps the parent's local orientation: its local node j is the centre of the smallest
        sub-entity of the parent (vertex, edge, facet, cell) that contains the parent's local
        vertices k and j - derived here from RefHex instead of a written-out table."""
        p, t, rd = self.p, self.t, self.refdom
        edge_node = self.t2e + p.shape[1]
        facet_node = self.t2f + p.shape[1] + self.edges.shape[1]
        cell_node = (np.arange(t.shape[1], dtype=np.int64)
                     + p.shape[1] + self.edges.shape[1] + self.facets.shape[1])
        newp = np.hstack((p,
                          .5 * np.sum(p[:, self.edges], axis=1),
                          .25 * np.sum(p[:, self.facets], axis=1),
                          .125 * np.sum(p[:, t], axis=1)))

        def centre(k, j):
            if k == j:
                return t[k]
            for i, e in enumerate(rd.edges):
                if {k, j} == set(e):
                    return edge_node[i]
            for i, f in enumerate(rd.facets):
                if {k, j} <= set(f):
                    return facet_node[i]
            return cell_node
        newt = np.hstack([np.vstack([centre(k, j) for j in range(8)]) for k in range(8)])
        return type(self)(newp, newt)


MeshTri1, MeshTet1, MeshHex1 = MeshTri, MeshTet, MeshHex
