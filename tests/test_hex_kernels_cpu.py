"""Hexahedral element-local kernels, run on the CPU from the shipped sources.

tests/host_block.py compiles ``local_hex_kernel`` (csrc/skb_local.cu) and
``local_hex_sf_kernel`` (csrc/skb_hex_sf.cu) with g++, one host thread per CUDA thread and a
``std::barrier`` for ``__syncthreads()``.  The scalar kernel follows the reference's operation
order: its output must equal the reference's element-local data (tests/golden, written by the
real reference) bit for bit, for ElementHex1 and ElementHex2.  The sum-factorised kernel
re-associates the quadrature sum: rtol 1e-12 on the local data and on the assembled CSR values,
bit-exact pattern, both output layouts identical, both CTA shapes identical, ragged last pass,
element subsets and the zero-Jacobian flag (mapping_isoparametric.py:195-196)."""
import ctypes as C

import numpy as np
import pytest

import host_block
import host_plan
import skfem_b200 as fem
from skfem_b200 import _lib, hex_sumfact
from cases import load
from product import mesh_from

RTOL = 1e-12
FORMS = {"laplace": _lib.FORM_LAPLACE, "mass": _lib.FORM_MASS}


def _space(basis, keep, tind=None):
    m = basis.mesh
    arrs = dict(p=np.ascontiguousarray(m.p), t=np.ascontiguousarray(m.t, dtype=np.int32),
                phi=np.ascontiguousarray(basis._phi), dphi=np.ascontiguousarray(basis._dphi),
                W=np.ascontiguousarray(basis.W), X=np.ascontiguousarray(basis.X),
                mphi=np.ascontiguousarray(basis._mphi), mdphi=np.ascontiguousarray(basis._mdphi))
    if tind is not None:
        arrs["tind"] = np.ascontiguousarray(tind, dtype=np.int32)
    keep.append(arrs)
    sp = _lib.SkbSpace()
    sp.dim, sp.nnodes, sp.mapping = 3, 8, _lib.SKB_MAP_ISO_HEX1
    sp.nbs, sp.ncomp, sp.nqp = basis.nbs, 1, basis.nqp
    sp.npts, sp.nel_total = m.p.shape[1], m.t.shape[1]
    sp.nel = m.t.shape[1] if tind is None else len(tind)
    sp.p, sp.t = arrs["p"].ctypes.data, arrs["t"].ctypes.data
    sp.tind = None if tind is None else arrs["tind"].ctypes.data
    for k in ("phi", "dphi", "W", "X", "mphi", "mdphi"):
        setattr(sp, k, arrs[k].ctypes.data)
    return sp


def _scalar(sp, form, nb, nel, bilinear=True, grid=3):
    out = np.full((nb * nb if bilinear else nb) * nel, np.nan)
    err = C.c_int(0)
    host_block.lib().host_local_hex(C.byref(sp), C.c_int(form), C.c_int(int(bilinear)),
                                    C.c_int(grid), C.c_void_p(out.ctypes.data), C.byref(err))
    return out, err.value


def _sumfact(sp, form, tab, nel, em=False, shape=0, grid=3):
    out = np.full(729 * nel, np.nan)
    err = C.c_int(0)
    p = lambda a: C.c_void_p(a.ctypes.data)
    host_block.lib().host_hex_sumfact(C.byref(sp), C.c_int(form), p(tab["qstride"]), p(tab["pp"]),
                                      p(tab["g"]), p(tab["bnode"]), p(tab["vtx"]), C.c_int(int(em)),
                                      C.c_int(shape), C.c_int(grid), p(out), C.byref(err))
    return out, err.value


@pytest.mark.parametrize("name,elem", [("hex1_tensor3", fem.ElementHex1),
                                       ("hex1_morphed3", fem.ElementHex1),
                                       ("hex2_morphed4", fem.ElementHex2),
                                       ("hex2_boxes3", fem.ElementHex2)])
def test_scalar_hex_kernel_matches_reference_bitwise(name, elem):
    g = load(name)
    basis = fem.Basis(mesh_from(g, "hex"), elem())
    keep = []
    sp = _space(basis, keep)
    nb, nel = basis.Nbfun, basis.nelems
    checked = 0
    for f, kid in FORMS.items():
        if f + "_local" in g.files:
            out, err = _scalar(sp, kid, nb, nel)
            assert err == 0
            assert np.array_equal(out, g[f + "_local"].reshape(-1)), (name, f)
            checked += 1
    if "unit_load_local" in g.files:
        out, err = _scalar(sp, _lib.LFORM_UNIT_LOAD, nb, nel, bilinear=False)
        assert np.array_equal(out, g["unit_load_local"].reshape(-1))
        checked += 1
    assert checked > 0


@pytest.mark.parametrize("name", ["hex2_tensor2", "hex2_morphed2", "hex2_morphed4",
                                  "hex2_boxes3"])
def test_sum_factorised_kernel_matches_reference(name):
    g = load(name)
    basis = fem.Basis(mesh_from(g, "hex"), fem.ElementHex2())
    tab = hex_sumfact.tables(basis)
    assert tab is not None
    keep = []
    sp = _space(basis, keep)
    nel, N = basis.nelems, basis.N
    edofs = np.ascontiguousarray(basis.element_dofs)
    for f, kid in FORMS.items():
        out, err = _sumfact(sp, kid, tab, nel)
        assert err == 0 and not np.isnan(out).any()
        if f + "_local" in g.files:
            loc = g[f + "_local"].reshape(-1)
            np.testing.assert_allclose(out, loc, rtol=RTOL, atol=RTOL * np.abs(loc).max())
        # element-major output: the same numbers, (nel, 27, 27); exactly symmetric local matrices
        em, _ = _sumfact(sp, kid, tab, nel, em=True)
        A3 = out.reshape(27, 27, nel)
        assert np.array_equal(em.reshape(nel, 27, 27), A3.transpose(2, 1, 0))
        assert np.array_equal(A3, A3.transpose(1, 0, 2))
        if kid == _lib.FORM_LAPLACE:                     # one CTA of 4 elements: same arithmetic
            assert np.array_equal(_sumfact(sp, kid, tab, nel, shape=1, grid=2)[0], out)
        # assembled through the plan / reduce kernels: the reference's CSR
        plan = host_plan.symbolic(out, edofs, edofs, nel, N, N, drop_zeros=True)
        assert np.array_equal(plan["indptr"], g[f + "_indptr"])
        assert np.array_equal(plan["indices"], g[f + "_indices"])
        data = host_plan.csr_reduce(out, plan)
        ref = g[f + "_data"]
        np.testing.assert_allclose(data, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())
        assert np.array_equal(host_plan.csr_reduce_em(em, nel, 27, 27, plan), data)


def test_sum_factorised_kernel_ragged_subset_and_zero_jacobian():
    x = np.linspace(0, 1, 4) ** 1.3
    m = fem.MeshHex.init_tensor(x, np.linspace(0, 2, 4), x)
    p = m.p.copy()
    p += 0.02 * np.sin(7 * p[[1, 2, 0]])
    m = fem.MeshHex(p, m.t)                              # 27 elements: ragged last pass
    basis = fem.Basis(m, fem.ElementHex2())
    tab = hex_sumfact.tables(basis)
    keep = []
    sp = _space(basis, keep)
    ref, err = _scalar(sp, _lib.FORM_LAPLACE, 27, 27)
    out, err2 = _sumfact(sp, _lib.FORM_LAPLACE, tab, 27)
    assert err == 0 and err2 == 0
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())
    sub = np.array([0, 5, 6, 13, 20, 26, 3])
    sps = _space(basis, keep, tind=sub)
    outs, _ = _sumfact(sps, _lib.FORM_LAPLACE, tab, len(sub), grid=2)
    assert np.array_equal(outs.reshape(729, len(sub)), out.reshape(729, 27)[:, sub])
    # a cell squeezed to zero volume raises the flag in both kernels
    p2 = p.copy()
    p2[2, m.t[:, 3]] = p2[2, m.t[0, 3]]
    b2 = fem.Basis(fem.MeshHex(p2, m.t), fem.ElementHex2())
    sp2 = _space(b2, keep)
    assert _scalar(sp2, _lib.FORM_LAPLACE, 27, 27)[1] == 1
    assert _sumfact(sp2, _lib.FORM_LAPLACE, tab, 27)[1] == 1
