"""The two plan builders of the generic path give the same arrays bit for bit.

``form.build_plan(method="rows")`` (csrc/skb_plan_rows.cu: row buckets, per-row sorts in shared
memory, own scans) against ``method="sort"`` (csrc/skb_plan.cu: one global radix sort), which
the golden tests pin to the reference's CSR (coo_data.py:27-36).  Covered: value-dependent
patterns with many exact zeros (Kuhn grids), rows sorted by one warp and by one CTA (vector P2:
up to ~2000 surviving entries per row), rows beyond the per-row capacity (falls back to the
radix sort), rectangular operators, 1-tensors, empty rows, element subsets."""
import numpy as np
import pytest
import torch

import skfem_b200 as fem
from skfem_b200 import form as F
from skfem_b200.models.elasticity import linear_elasticity
from skfem_b200.models.poisson import laplace, mass, unit_load
from cases import LAME, load
from product import mesh_from

pytestmark = pytest.mark.gpu


def _plans(form, ub, vb=None, drop_zeros=True):
    vb = ub if vb is None else vb
    if isinstance(form, F.LinearForm):
        args = (ub._dev()["edofs"], None, ub.nelems, (ub.N,), None)
        kw = dict(drop_zeros=False)
    else:
        local = form._local(ub, None if vb is ub else vb)
        args = (vb._dev()["edofs"], ub._dev()["edofs"], ub.nelems, (vb.N, ub.N), local)
        kw = dict(drop_zeros=drop_zeros)
    a = F.build_plan(*args, method="rows", **kw)
    b = F.build_plan(*args, method="sort", **kw)
    return a, b


def _same(a, b):
    assert a.nnz == b.nnz and a.nkeep == b.nkeep and a.shape == b.shape
    for name in ("indptr", "indices", "segptr"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.equal(a.perm[:a.nkeep], b.perm[:b.nkeep])


@pytest.mark.parametrize("name,refdom,elem,vector", [
    ("tet_p1_tensor6", "tet", fem.ElementTetP1, False),
    ("tet_p1_ball2", "tet", fem.ElementTetP1, False),
    ("tet_p2_tensor3", "tet", fem.ElementTetP2, False),
    ("tet_vp2_elasticity2", "tet", fem.ElementTetP2, True),
    ("tri_p2_morphed3", "tri", fem.ElementTriP2, False),
    ("hex1_morphed3", "hex", fem.ElementHex1, False),
    ("hex2_morphed4", "hex", fem.ElementHex2, False),
])
def test_rows_plan_equals_sort_plan_on_fixtures(name, refdom, elem, vector):
    g = load(name)
    e = fem.ElementVector(elem()) if vector else elem()
    b = fem.Basis(mesh_from(g, refdom), e)
    forms = [linear_elasticity(*LAME)] if vector else [laplace, mass, unit_load]
    for f in forms:
        for dz in (True, False):
            _same(*_plans(f, b, drop_zeros=dz))


def test_rows_plan_long_rows_and_zero_dropping():
    """Vector P2 on a Kuhn grid: rows with more than 512 surviving entries go to the CTA-wide
    sort, and the grid's exact zeros make the pattern value dependent."""
    x = np.linspace(0, 1, 6)
    m = fem.MeshTet.init_tensor(x, x, x)
    b = fem.Basis(m, fem.ElementVector(fem.ElementTetP2()))
    a, s = _plans(linear_elasticity(*LAME), b)
    _same(a, s)
    rows = (a.indptr[1:] - a.indptr[:-1]).max().item()
    per_row = torch.zeros(b.N, dtype=torch.int64, device=a.perm.device)
    seg = (a.segptr[1:].long() - a.segptr[:-1].long())
    per_row.scatter_add_(0, torch.repeat_interleave(
        torch.arange(b.N, device=seg.device), (a.indptr[1:] - a.indptr[:-1]).long()), seg)
    # all three sort kernels were exercised: registers (<= 128 entries per row), one warp in
    # shared memory (<= 512), one CTA
    assert per_row.min().item() <= 128 and per_row.max().item() > 128 and rows > 60
    _same(*_plans(linear_elasticity(*LAME), b, drop_zeros=False))
    full = 30 * torch.zeros(b.N, dtype=torch.int64).scatter_add_(
        0, torch.from_numpy(b.element_dofs.astype(np.int64).reshape(-1)),
        torch.ones(b.element_dofs.size, dtype=torch.int64))
    assert full.max().item() > 512 and full.min().item() <= 128
    assert a.nkeep < a.ncoo                           # exact zeros were dropped
    # and the assembled matrix through the public API is the radix-sort one
    F.set_options(plan_method="rows")
    A = linear_elasticity(*LAME).assemble(b)
    F.set_options(plan_method="sort")
    try:
        B = linear_elasticity(*LAME).assemble(fem.Basis(m, fem.ElementVector(fem.ElementTetP2())))
    finally:
        F.set_options(plan_method="rows")
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert np.array_equal(A.data, B.data)


def test_rows_plan_falls_back_beyond_the_row_capacity():
    """A fan of 2600 tetrahedra around one edge: the two hub vertices collect 4 x 2600 = 10400
    entries each, more than a CTA sorts in shared memory - build_plan must still return the
    right plan (radix-sort fallback)."""
    n = 2600
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    p = np.vstack([np.r_[0.0, 0.0, np.cos(ang)], np.r_[0.0, 0.0, np.sin(ang)],
                   np.r_[0.0, 1.0, np.zeros(n)]])
    t = np.vstack([np.zeros(n, dtype=np.int64), np.ones(n, dtype=np.int64),
                   2 + np.arange(n), 2 + (np.arange(n) + 1) % n])
    m = fem.MeshTet(p, t)
    b = fem.Basis(m, fem.ElementTetP1())
    a, s = _plans(mass, b)
    _same(a, s)
    assert int((a.segptr[1:].long() - a.segptr[:-1].long()).max()) == n


def test_rows_plan_rectangular_subset_and_empty_rows():
    g = load("tet_p1_morphed5")
    m = mesh_from(g, "tet")
    sub = np.arange(0, m.nelements, 7)                # most vertices see no element
    ub = fem.Basis(m, fem.ElementTetP2(), elements=sub)
    vb = fem.Basis(m, fem.ElementTetP1(), elements=sub, intorder=4)
    ub4 = fem.Basis(m, fem.ElementTetP2(), elements=sub, intorder=4)
    form = fem.BilinearForm(lambda u, v, w: u * v)
    a, s = _plans(form, ub4, vb)
    _same(a, s)
    assert a.shape == (vb.N, ub4.N) and int((a.indptr[1:] == a.indptr[:-1]).sum()) > 0
    _same(*_plans(unit_load, ub))
