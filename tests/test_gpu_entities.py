"""Mesh.build_entities on the GPU: edges of any mesh and the facets of
tetrahedra numbered by the library's own row-bucket sort (csrc/skb_plan_rows.cu through
skb_entity_masks / skb_plan_slot_of_entry), quadrilateral facets by packed two-level keys -
entities and incidences identical to the host path (np.sort + np.unique, mesh/mesh.py:1065-1082,
which tests/test_oracle_golden.py pins to the reference's numbering)."""
import numpy as np
import pytest

import skfem_b200 as fem
from skfem_b200 import mesh as M

pytestmark = pytest.mark.gpu


def _host(monkeypatch, t, indices, sort=True):
    monkeypatch.setattr(M, "_cuda_ready", lambda: False)
    try:
        return M.Mesh.build_entities(t, indices, sort=sort)
    finally:
        monkeypatch.undo()


def _check(monkeypatch, mesh):
    rd = mesh.refdom
    for indices, sort in ((rd.edges, True), (rd.facets, mesh._sort_facets)):
        if indices is None:
            continue
        ent_h, inc_h = _host(monkeypatch, mesh.t, indices, sort)
        ent_d, inc_d = M.Mesh.build_entities(mesh.t, indices, sort=sort)
        assert ent_d.dtype == ent_h.dtype and ent_d.shape == ent_h.shape
        assert np.array_equal(ent_d, ent_h)
        assert np.array_equal(inc_d, inc_h)


def test_tet_entities_structured_and_shuffled(monkeypatch):
    x = np.linspace(0, 1, 24)
    m = fem.MeshTet.init_tensor(x, x, x)                  # 73 002 tets
    assert m.nelements >= (1 << 16)
    _check(monkeypatch, m)
    rng = np.random.default_rng(1)
    perm = rng.permutation(m.p.shape[1])                  # unstructured vertex numbering
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    m2 = fem.MeshTet(m.p[:, perm], inv[m.t].astype(m.t.dtype))
    _check(monkeypatch, m2)
    # and through the public surface: P2 numbering == host numbering
    monkeypatch.setattr(M, "_cuda_ready", lambda: False)
    bh = fem.Basis(fem.MeshTet(m2.p, m2.t), fem.ElementTetP2())
    monkeypatch.undo()
    bd = fem.Basis(fem.MeshTet(m2.p, m2.t), fem.ElementTetP2())
    assert np.array_equal(bd.element_dofs, bh.element_dofs) and bd.N == bh.N


def test_tri_and_hex_entities(monkeypatch):
    x = np.linspace(0, 1, 202)
    _check(monkeypatch, fem.MeshTri.init_tensor(x, x))    # 80 802 triangles
    xh = np.linspace(0, 1, 42)
    mh = fem.MeshHex.init_tensor(xh, xh, xh)              # 68 921 hexes: quad facets, 12 edges
    _check(monkeypatch, mh)
