"""Global degree-of-freedom numbering.

Bit-exact restatement of the layout produced by the reference's
``Dofs.__init__`` (skfem/assembly/dofs.py:264-334): vertex DOFs first
(``nd*vertex + component``), then edge, facet and interior blocks; the
``element_dofs`` rows follow local vertices, local edges (``t2e`` rows), local
facets (``t2f`` rows), interior.  Everything is int32.

Host side: the numbering is built once per (mesh, element) from the mesh
topology and uploaded; the kernels only ever read ``element_dofs``.
"""
from __future__ import annotations

import numpy as np


def _block(ndofs_per_entity, nentities, offset):
    """(ndofs, nentities) table numbered entity-major, starting at offset."""
    tab = np.arange(ndofs_per_entity * nentities, dtype=np.int32)
    return tab.reshape((ndofs_per_entity, nentities), order='F') + np.int32(offset)


class DofsView:
    """A subset of a :class:`Dofs` numbering - what ``basis.get_dofs(...)`` returns (the
    reference's ``DofsView``, dofs.py:17-262).  It remembers *which* vertices / facets / edges /
    cells were selected and which DOF rows of each are kept, so that the set can be narrowed
    by name (``keep`` / ``drop`` / ``all('u^1')``) or split by entity type (``nodal`` ...).
    Anything numpy-like (``np.asarray``, fancy indexing, ``len``, iteration) sees the sorted
    unique DOF indices, an int32 array."""

    KINDS = ("nodal", "facet", "edge", "interior")      # the order of ``element.dofnames``

    def __init__(self, dofs, entities, rows=None, doflocs=None):
        self.dofs = dofs
        self.entities = {k: np.asarray(entities.get(k, np.zeros(0, np.int32))) for k in self.KINDS}
        self.rows = rows if rows is not None else {k: list(range(self._table(k).shape[0]))
                                                   for k in self.KINDS}
        self._doflocs = doflocs                          # callable -> (dim, N) array, or None

    # -- tables ------------------------------------------------------------------
    def _table(self, kind):
        return getattr(self.dofs, kind + "_dofs")

    def _name_offset(self, kind):
        return sum(self._table(k).shape[0] for k in self.KINDS[:self.KINDS.index(kind)])

    def _selected(self, kind):
        tab = self._table(kind)
        if tab.size == 0 or len(self.rows[kind]) == 0:
            return np.zeros((0, 0), dtype=np.int32)
        return tab[self.rows[kind]][:, self.entities[kind]]

    def _by_name(self, kind):
        """{dof name: indices}; rows that share a name are concatenated row by row."""
        if self._table(kind).size == 0:
            return {}
        names = self.dofs.element.dofnames
        off, out = self._name_offset(kind), {}
        sel = self._selected(kind)
        for i, r in enumerate(self.rows[kind]):
            row = sel[i] if sel.size else np.zeros(0, dtype=np.int32)
            out.setdefault(names[off + r], []).append(row)
        return {k: np.concatenate(v).astype(np.int32) for k, v in out.items()}

    nodal = property(lambda self: self._by_name("nodal"))
    facet = property(lambda self: self._by_name("facet"))
    edge = property(lambda self: self._by_name("edge"))
    interior = property(lambda self: self._by_name("interior"))

    # -- the flat view -------------------------------------------------------------
    def flatten(self):
        parts = [self._selected(k).reshape(-1) for k in self.KINDS]
        return np.unique(np.concatenate(parts)).astype(np.int32)

    def all(self, key=None):
        return self.flatten() if key is None else self.keep(key).flatten()

    def __array__(self, dtype=None, copy=None):
        flat = self.flatten()
        return flat if dtype is None else flat.astype(dtype)

    def __len__(self):
        return len(self.flatten())

    def __iter__(self):
        return iter(self.flatten())

    def __getitem__(self, ix):
        return self.flatten()[ix]

    def __repr__(self):
        counts = ", ".join("{} {}".format(len(np.unique(self._selected(k))), k)
                           for k in self.KINDS if self._selected(k).size)
        return "<skfem_b200 DofsView({}): {} DOFs ({})>".format(
            type(self.dofs.element).__name__, len(self), counts)

    # -- narrowing by DOF name -----------------------------------------------------
    def _filtered(self, dofnames, keep):
        if isinstance(dofnames, str):
            dofnames = [dofnames]
        names = self.dofs.element.dofnames
        rows = {k: [r for r in self.rows[k]
                    if (names[self._name_offset(k) + r] in dofnames) == keep]
                for k in self.KINDS}
        return DofsView(self.dofs, self.entities, rows, self._doflocs)

    def keep(self, dofnames):
        """Only the DOFs with the given names, e.g. ``['u^1']``."""
        return self._filtered(dofnames, True)

    def drop(self, dofnames):
        """Everything but the DOFs with the given names."""
        return self._filtered(dofnames, False)

    def sort(self, sorting=None):
        """The DOF indices ordered by ``sorting(doflocs)`` (default: the coordinate sum)."""
        if self._doflocs is None:
            raise NotImplementedError("DofsView.sort needs the DOF locations of a basis")
        flat = self.flatten()
        x = self._doflocs()[:, flat]
        return flat[np.argsort(sum(x) if sorting is None else sorting(x))]

    def __or__(self, other):
        ents = {k: np.union1d(self.entities[k], other.entities[k]) for k in self.KINDS}
        return DofsView(self.dofs, ents, self.rows, self._doflocs)

    __add__ = __or__


class _LazyBlock:
    """Descriptor: a ``_block`` table built on first access (a million-entry arange per
    Basis is a millisecond of the cold path that most assemblies never look at)."""

    def __init__(self, name):
        self.name = name

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        tab = obj.__dict__.get(self.name)
        if tab is None:
            nd, n, off = obj._blocks[self.name]
            tab = _block(nd, n, off) if n else np.empty((0, 0), dtype=np.int32)
            obj.__dict__[self.name] = tab
        return tab


class Dofs:
    nodal_dofs = _LazyBlock("nodal_dofs")
    edge_dofs = _LazyBlock("edge_dofs")
    facet_dofs = _LazyBlock("facet_dofs")
    interior_dofs = _LazyBlock("interior_dofs")

    def __init__(self, topo, element, offset=0):
        self.topo = topo
        self.element = element
        nel = topo.nelements
        three_d = element.dim == 3
        # (dofs per entity, entities, first index) of the four blocks; the tables themselves
        # are built lazily, entity-major like the reference (dofs.py:264-334)
        self._blocks = {}
        off0 = offset
        self._blocks["nodal_dofs"] = (element.nodal_dofs, topo.nvertices, offset)
        offset += element.nodal_dofs * topo.nvertices
        nedge = element.edge_dofs if three_d else 0
        self._blocks["edge_dofs"] = (nedge, topo.nedges if nedge else 0, offset)
        offset += nedge * (topo.nedges if nedge else 0)
        nfac = element.facet_dofs
        self._blocks["facet_dofs"] = (nfac, topo.nfacets if nfac else 0, offset)
        offset += nfac * (topo.nfacets if nfac else 0)
        self._blocks["interior_dofs"] = (element.interior_dofs, nel, offset)
        if element.interior_dofs == 0:      # the reference's (0, nel) table
            self.__dict__["interior_dofs"] = np.empty((0, nel), dtype=np.int32)

        # every row of element_dofs is  first + c + nd * entity(e)  (entity-major blocks), i.e.
        # integer arithmetic on t / t2e / t2f; for one DOF per vertex element_dofs IS t (no
        # copy, no extra upload)
        nd = element.nodal_dofs
        self.nodal_is_t = False
        self._edofs_dev = None
        rows = []                                   # (source, row of the source, mul, add)
        for k in range(topo.t.shape[0]):
            rows += [("t", k, nd, off0 + c) for c in range(nd)]
        if nedge:
            for k in range(len(topo.refdom.edges)):       # == t2e.shape[0], without fetching it
                rows += [("t2e", k, nedge, self._blocks["edge_dofs"][2] + c) for c in range(nedge)]
        if element.dim >= 2 and nfac:
            for k in range(len(topo.refdom.facets)):      # == t2f.shape[0]
                rows += [("t2f", k, nfac, self._blocks["facet_dofs"][2] + c) for c in range(nfac)]
        ni = element.interior_dofs
        rows += [(None, 0, ni, self._blocks["interior_dofs"][2] + c) for c in range(ni)]
        if nd == 1 and off0 == 0 and len(rows) == topo.t.shape[0]:
            self.nodal_is_t = True
            self.element_dofs = topo.t
        else:
            self.element_dofs = self._element_dofs_device(topo, rows)
            if self.element_dofs is None:
                src = {"t": topo.t, "t2e": topo.t2e if nedge else None,
                       "t2f": topo.t2f if (element.dim >= 2 and nfac) else None}
                e = np.arange(nel, dtype=np.int32)
                self.element_dofs = np.ascontiguousarray(np.vstack(
                    [np.int32(add) + np.int32(mul) * (e if s is None else src[s][r])
                     for s, r, mul, add in rows]), dtype=np.int32)
        # == max(element_dofs) + 1: every vertex/edge/facet/cell is referenced
        self.N = int(offset + element.interior_dofs * nel)

    def _element_dofs_device(self, topo, rows):
        """element_dofs by csrc/skb_mesh.cu when the mesh (and the incidences needed) already
        live on the GPU: the host copy comes back through pinned memory, the device copy is
        kept for the basis (no re-upload).  None: build on the host."""
        devs = getattr(topo, "_dev", None)
        if not devs or topo.nelements < (1 << 16):
            return None
        import ctypes as C
        import torch
        from . import _lib
        key, (p_dev, t_dev) = next(iter(devs.items()))
        src = {"t": t_dev, "t2e": getattr(topo, "_t2e_dev", None),
               "t2f": getattr(topo, "_t2f_dev", None)}
        if any(s is not None and src[s] is None for s, _, _, _ in rows):
            return None
        nel = int(t_dev.shape[1])
        desc = np.zeros(len(rows), dtype=[("src", np.int64), ("mul", np.int32), ("add", np.int32)])
        for i, (s, r, mul, add) in enumerate(rows):
            desc[i] = (0 if s is None else src[s][r].data_ptr(), mul, add)
        desc_dev = torch.from_numpy(desc.view(np.uint8)).to(t_dev.device)
        out = torch.empty((len(rows), nel), dtype=torch.int32, device=t_dev.device)
        code = _lib.lib().skb_element_dofs(desc_dev.data_ptr(), len(rows), nel, out.data_ptr(),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(code, "skb_element_dofs")
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._edofs_dev = (key, out)
        self._edofs_pinned = host
        return host.numpy()

    def on_facets(self, facets):
        """All DOFs attached to the vertices / edges / facets of the given facets,
        sorted - what ``get_facet_dofs(facets).all()`` returns in the reference
        (dofs.py:618-663, 90-107)."""
        m = self.topo
        facets = np.asarray(facets, dtype=np.int64)
        out = [self.nodal_dofs[:, np.unique(m.facets[:, facets])].flatten()]
        if self.edge_dofs.size:
            out.append(self.edge_dofs[:, m.facet_edges(facets)].flatten())
        if self.facet_dofs.size:
            out.append(self.facet_dofs[:, facets].flatten())
        return np.unique(np.concatenate(out)).astype(np.int32)

    def boundary(self):
        """All DOFs attached to boundary vertices / edges / facets."""
        return self.on_facets(self.topo.boundary_facets())

    # -- views (dofs.py:536-663): which entities carry the selected DOFs ---------------
    def _view(self, entities, skip, doflocs):
        el = self.element
        empty = np.zeros(0, dtype=np.int32)
        ents = {"nodal": entities.get("nodal", empty) if el.nodal_dofs > 0 else empty,
                "edge": entities.get("edge", empty) if self.edge_dofs.size else empty,
                "facet": entities.get("facet", empty) if el.facet_dofs > 0 else empty,
                "interior": entities.get("interior", empty)}
        view = DofsView(self, ents, doflocs=doflocs)
        return view.drop(skip) if skip else view

    def facet_view(self, facets, skip=None, doflocs=None):
        m = self.topo
        facets = np.asarray(facets, dtype=np.int64)
        ents = {"facet": facets}
        if self.element.nodal_dofs > 0:
            ents["nodal"] = np.unique(m.facets[:, facets])
        if self.edge_dofs.size:
            ents["edge"] = m.facet_edges(facets)
        return self._view(ents, skip, doflocs)

    def element_view(self, elements, skip=None, doflocs=None):
        m = self.topo
        elements = np.asarray(elements, dtype=np.int64)
        ents = {"interior": elements}
        if self.element.nodal_dofs > 0:
            ents["nodal"] = np.unique(m.t[:, elements])
        if self.edge_dofs.size:
            ents["edge"] = np.unique(m.t2e[:, elements])
        if self.element.facet_dofs > 0:
            ents["facet"] = np.unique(m.t2f[:, elements])
        return self._view(ents, skip, doflocs)

    def vertex_view(self, nodes, skip=None, doflocs=None):
        return self._view({"nodal": np.asarray(nodes, dtype=np.int64)}, skip, doflocs)
