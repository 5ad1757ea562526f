"""Element-local kernels of affine meshes, run on the CPU from the shipped source.

tests/host_local.py compiles ``local_affine_kernel`` / ``local_affine_cached_kernel``
(csrc/skb_local.cu + csrc/skb_common.cuh) with g++ as a single-thread grid; here their output
must equal, bit for bit, the reference's element-local data ``Form.elemental(basis).data`` stored
in tests/golden (written by the real reference) for every library form and every affine
fixture - the same ``array_equal`` bar the GPU parity tests apply to the device build."""
import ctypes as C

import numpy as np
import pytest

import host_local
import skfem_b200 as fem
from skfem_b200 import _lib
from cases import CASES, LAME, load
from product import mesh_from, element_from

NATIVE = {"laplace": (_lib.FORM_LAPLACE, False), "mass": (_lib.FORM_MASS, None),
          "vector_laplace": (_lib.FORM_VECTOR_LAPLACE, True),
          "elasticity": (_lib.FORM_ELASTICITY, True)}
AFFINE = [n for n, c in CASES.items() if c[0] in ("tri", "tet")]


def _space(basis, keep):
    m = basis.mesh
    arrs = dict(p=np.ascontiguousarray(m.p), t=np.ascontiguousarray(m.t, dtype=np.int32),
                phi=np.ascontiguousarray(basis._phi), dphi=np.ascontiguousarray(basis._dphi),
                W=np.ascontiguousarray(basis.W), X=np.ascontiguousarray(basis.X))
    keep.append(arrs)
    sp = _lib.SkbSpace()
    sp.dim, sp.nnodes, sp.mapping = m.dim(), m.t.shape[0], _lib.SKB_MAP_AFFINE
    sp.nbs, sp.ncomp, sp.nqp = basis.nbs, basis.ncomp, basis.nqp
    sp.npts, sp.nel_total, sp.nel = m.p.shape[1], m.t.shape[1], m.t.shape[1]
    sp.p, sp.t, sp.tind = arrs["p"].ctypes.data, arrs["t"].ctypes.data, None
    sp.phi, sp.dphi, sp.W, sp.X = (arrs[k].ctypes.data for k in ("phi", "dphi", "W", "X"))
    sp.mdphi = sp.mphi = None
    return sp


@pytest.mark.parametrize("name", AFFINE)
def test_local_kernels_match_reference_bitwise(name):
    refdom, ename, vector, bil, lin, has_local = CASES[name]
    g = load(name)
    basis = fem.Basis(mesh_from(g, refdom), element_from(ename, vector))
    keep = []
    sp = _space(basis, keep)
    lib = host_local.lib()
    nb, nel = basis.Nbfun, basis.nelems
    checked = 0
    for f in bil:
        if f not in NATIVE or not has_local:
            continue
        kid, _ = NATIVE[f]
        lam, two_mu = (LAME[0], 2. * LAME[1]) if f == "elasticity" else (1.0, 2.0)
        # vector elements: dense, cached and cached with a compile-time rule size; scalar: dense
        # and register-cached symmetric
        variants = [0, 1, 3] if vector else [0, 2]
        for cached in variants:
            out = np.full(nb * nb * nel, np.nan)
            rc = lib.host_local_affine(C.byref(sp), C.c_int(kid), C.c_double(lam),
                                       C.c_double(two_mu), C.c_void_p(out.ctypes.data),
                                       C.c_int(1), C.c_int(cached))
            assert rc == 0
            assert np.array_equal(out, g[f + "_local"]), (name, f, cached)
            checked += 1
    if "unit_load" in lin and not vector:
        out = np.full(nb * nel, np.nan)
        lib.host_local_affine(C.byref(sp), C.c_int(_lib.LFORM_UNIT_LOAD), C.c_double(1.0),
                              C.c_double(2.0), C.c_void_p(out.ctypes.data), C.c_int(0), C.c_int(0))
        vec = np.zeros(basis.N)
        np.add.at(vec, basis.element_dofs.reshape(-1), out)     # sequential, COO order (A.10)
        assert np.array_equal(vec, g["unit_load_vec"]), name
        checked += 1
    assert checked > 0


@pytest.mark.parametrize("name", AFFINE)
def test_element_major_kernels_and_reduce_on_the_cpu(name):
    """The element-major variants of the fixed-rule kernels (skb_local_bilinear_em) write the
    reference's element-local data bit for bit, transposed to (nel, Nbv, Nbu), and
    ``csr_reduce_em_kernel`` turns them into the same CSR values as ``csr_reduce_kernel`` does
    from the reference layout - from the shipped sources, on the host."""
    import host_plan
    refdom, ename, vector, bil, lin, has_local = CASES[name]
    g = load(name)
    basis = fem.Basis(mesh_from(g, refdom), element_from(ename, vector))
    keep = []
    sp = _space(basis, keep)
    lib = host_local.lib()
    nb, nel, N = basis.Nbfun, basis.nelems, basis.N
    edofs = np.ascontiguousarray(basis.element_dofs)
    checked = 0
    for f in bil:
        if f not in NATIVE:
            continue
        kid, _ = NATIVE[f]
        lam, two_mu = (LAME[0], 2. * LAME[1]) if f == "elasticity" else (1.0, 2.0)
        outs = {}
        for cached in ((3, 5) if vector else (2, 4)):
            out = np.full(nb * nb * nel, np.nan)
            rc = lib.host_local_affine(C.byref(sp), C.c_int(kid), C.c_double(lam),
                                       C.c_double(two_mu), C.c_void_p(out.ctypes.data),
                                       C.c_int(1), C.c_int(cached))
            assert rc == 0, (name, f, cached)
            outs[cached] = out
        local, em = outs[3 if vector else 2], outs[5 if vector else 4]
        if has_local:
            assert np.array_equal(local, g[f + "_local"])
        # em[e][i][j] == local[j][i][e]
        assert np.array_equal(em.reshape(nel, nb, nb), local.reshape(nb, nb, nel).transpose(2, 1, 0))
        plan = host_plan.symbolic(local, edofs, edofs, nel, N, N, drop_zeros=True)
        data = host_plan.csr_reduce(local, plan)
        assert np.array_equal(host_plan.csr_reduce_em(em, nel, nb, nb, plan), data), (name, f)
        _check_csr(plan, data, g, f)
        checked += 1
    assert checked > 0


def _check_csr(plan, data, g, f):
    assert np.array_equal(plan["indptr"], g[f + "_indptr"]), f
    assert np.array_equal(plan["indices"], g[f + "_indices"]), f
    ref = g[f + "_data"]
    np.testing.assert_allclose(data, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("name", AFFINE)
def test_generic_pipeline_on_the_cpu_matches_reference_csr(name):
    """local kernel -> plan kernels (+ stable sort / scan) -> csr_reduce / vec_reduce, all from the
    shipped sources compiled for the host (tests/host_local.py, tests/host_plan.py): indptr and
    indices bit-exact, values within rtol 1e-12, load vectors bit-exact - the bar of
    BASELINE.json's north_star, met without a GPU."""
    import host_plan
    refdom, ename, vector, bil, lin, has_local = CASES[name]
    g = load(name)
    basis = fem.Basis(mesh_from(g, refdom), element_from(ename, vector))
    keep = []
    sp = _space(basis, keep)
    lib = host_local.lib()
    nb, nel, N = basis.Nbfun, basis.nelems, basis.N
    edofs = np.ascontiguousarray(basis.element_dofs)
    for f in bil:
        if f not in NATIVE:
            continue
        lam, two_mu = (LAME[0], 2. * LAME[1]) if f == "elasticity" else (1.0, 2.0)
        local = np.full(nb * nb * nel, np.nan)
        lib.host_local_affine(C.byref(sp), C.c_int(NATIVE[f][0]), C.c_double(lam),
                              C.c_double(two_mu), C.c_void_p(local.ctypes.data), C.c_int(1),
                              C.c_int(1 if vector else 0))
        plan = host_plan.symbolic(local, edofs, edofs, nel, N, N, drop_zeros=True)
        _check_csr(plan, host_plan.csr_reduce(local, plan), g, f)
    if "unit_load" in lin and not vector:
        local = np.full(nb * nel, np.nan)
        lib.host_local_affine(C.byref(sp), C.c_int(_lib.LFORM_UNIT_LOAD), C.c_double(1.0),
                              C.c_double(2.0), C.c_void_p(local.ctypes.data), C.c_int(0), C.c_int(0))
        plan = host_plan.symbolic(None, edofs, None, nel, N, 1, drop_zeros=False)
        assert np.array_equal(host_plan.vec_reduce(local, plan, N), g["unit_load_vec"])
