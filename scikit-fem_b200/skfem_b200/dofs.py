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


class Dofs:

    def __init__(self, topo, element, offset=0):
        self.topo = topo
        self.element = element
        nel = topo.nelements
        three_d = element.dim == 3

        self.nodal_dofs = _block(element.nodal_dofs, topo.nvertices, offset)
        offset += self.nodal_dofs.size

        if three_d and element.edge_dofs > 0:
            self.edge_dofs = _block(element.edge_dofs, topo.nedges, offset)
            offset += self.edge_dofs.size
        else:
            self.edge_dofs = np.empty((0, 0), dtype=np.int32)

        if element.facet_dofs > 0:
            self.facet_dofs = _block(element.facet_dofs, topo.nfacets, offset)
            offset += self.facet_dofs.size
        else:
            self.facet_dofs = np.empty((0, 0), dtype=np.int32)

        self.interior_dofs = _block(element.interior_dofs, nel, offset)

        # nodal rows: nodal_dofs[c, v] == nd*v + c + offset0, so the gather
        # nodal_dofs[:, t[k]] is plain integer arithmetic on t (and for one
        # DOF per vertex element_dofs IS t: no copy, no extra upload)
        nd, off0 = element.nodal_dofs, np.int32(self.nodal_dofs[0, 0]) if self.nodal_dofs.size else 0
        self.nodal_is_t = False
        if nd == 1 and off0 == 0:
            parts = [topo.t]
            self.nodal_is_t = True
        else:
            parts = [np.int32(nd) * topo.t[k][None, :]
                     + (np.arange(nd, dtype=np.int32) + off0)[:, None]
                     for k in range(topo.t.shape[0])] if nd > 0 else []
        if self.edge_dofs.size:
            parts += [self.edge_dofs[:, topo.t2e[k]] for k in range(topo.t2e.shape[0])]
        if element.dim >= 2 and self.facet_dofs.size:
            parts += [self.facet_dofs[:, topo.t2f[k]] for k in range(topo.t2f.shape[0])]
        if self.interior_dofs.size:
            parts.append(self.interior_dofs)
        if len(parts) == 1 and self.nodal_is_t:
            self.element_dofs = topo.t
        else:
            self.nodal_is_t = False
            self.element_dofs = np.ascontiguousarray(np.vstack(parts), dtype=np.int32)
        # == max(element_dofs) + 1: every vertex/edge/facet/cell is referenced
        self.N = int(offset + self.interior_dofs.size)

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
