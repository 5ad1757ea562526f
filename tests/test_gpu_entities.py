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


def test_init_tensor_on_the_device_equals_the_host_generator(monkeypatch):
    """MeshTet / MeshHex.init_tensor of large grids run on the GPU (csrc/skb_mesh.cu): p and t
    bit for bit the host generator's (unsorted, non-uniform inputs included), and the device
    copies are adopted as the mesh's device arrays."""
    import torch
    rng = np.random.default_rng(7)
    x, y, z = rng.random(31), rng.random(27), rng.random(29)
    for cls in (fem.MeshTet, fem.MeshHex):
        if cls is fem.MeshHex:
            x, y, z = rng.random(45), rng.random(41), rng.random(43)
        md = cls.init_tensor(x, y, z)
        assert md._dev, "device path not taken"
        monkeypatch.setattr(M, "_cuda_ready", lambda: False)
        mh = cls.init_tensor(x, y, z)
        monkeypatch.undo()
        assert not mh._dev
        assert md.p.dtype == mh.p.dtype and md.t.dtype == mh.t.dtype
        assert np.array_equal(md.p, mh.p) and np.array_equal(md.t, mh.t)
        pd, td = next(iter(md._dev.values()))
        assert torch.equal(pd.cpu(), torch.from_numpy(mh.p))
        assert torch.equal(td.cpu(), torch.from_numpy(mh.t))
    # the generated mesh assembles like any other
    from skfem_b200.models.poisson import laplace
    g = np.linspace(0, 1, 24)
    A = laplace.assemble(fem.Basis(fem.MeshTet.init_tensor(g, g, g), fem.ElementTetP1()))
    assert A.nnz == 24 ** 3 + 6 * 23 * 24 ** 2
