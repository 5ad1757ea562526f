"""FacetBasis kernels (csrc/skb_facet.cu) run on the CPU from the shipped source
(tests/host_facet.py) against the reference's FacetBasis (tests/golden/facet_*.npz, written by
the real reference): global points, normals, dx and every basis function with its gradient at the
facet quadrature points, bit for bit."""
import ctypes as C

import numpy as np
import pytest

import host_facet
import skfem_b200 as fem
from skfem_b200 import _lib
from cases import load


def _p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("name,M,E", [
    ("facet_tri_p1", "MeshTri", "ElementTriP1"), ("facet_tri_p2", "MeshTri", "ElementTriP2"),
    ("facet_tet_p1", "MeshTet", "ElementTetP1"), ("facet_tet_p2", "MeshTet", "ElementTetP2")])
def test_facet_geometry_and_basis_match_reference_bitwise(name, M, E):
    g = load(name)
    m = getattr(fem, M)(g["p"], g["t"])
    fb = fem.FacetBasis(m, getattr(fem, E)())
    lib = host_facet.lib()
    dim, nf, nqp = m.dim(), fb.nelems, fb.nqp
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)          # noqa: E731
    p, t, facets = np.ascontiguousarray(m.p), i32(m.t), i32(m.facets)
    find, tind, tind_n = i32(fb.find), i32(fb.tind), i32(fb.tind_normals)
    lfacet = i32(np.argmax(m.t2f[:, fb.tind_normals] == np.asarray(fb.find)[None, :], axis=0))
    Xb, Wb = np.ascontiguousarray(fb.X), np.ascontiguousarray(fb.W)
    sp = _lib.SkbSpace()
    sp.dim, sp.nnodes, sp.mapping = dim, m.t.shape[0], _lib.SKB_MAP_AFFINE
    sp.nbs, sp.ncomp, sp.nqp = fb.nbs, fb.ncomp, nqp
    sp.npts, sp.nel_total, sp.nel = m.p.shape[1], m.t.shape[1], nf
    sp.p, sp.t, sp.tind = p.ctypes.data, t.ctypes.data, tind.ctypes.data
    x, Y, nrm = (np.full((dim, nf, nqp), np.nan) for _ in range(3))
    dx, detabs = np.full((nf, nqp), np.nan), np.full((nf, nqp), np.nan)
    lib.host_facet_geometry(C.byref(sp), _p(facets), C.c_int64(facets.shape[1]), _p(find),
                            _p(tind), _p(tind_n), _p(lfacet), C.c_int64(nf), _p(Xb), _p(Wb),
                            C.c_int(nqp), _p(x), _p(Y), _p(dx), _p(nrm), _p(detabs))
    assert np.array_equal(x, g["x"]) and np.array_equal(dx, g["dx"])
    assert np.array_equal(nrm, g["normals"])
    coef, expo, nterm = (np.ascontiguousarray(a) for a in fb._tables)
    for b in range(fb.nbs):
        val, grad = np.full((nf, nqp), np.nan), np.full((dim, nf, nqp), np.nan)
        lib.host_facet_basis(C.byref(sp), _p(tind), C.c_int64(nf), C.c_int(nqp), _p(Y), _p(coef),
                             _p(expo), _p(nterm), C.c_int(b), _p(val), _p(grad))
        assert np.array_equal(val, g["phi"][b]), (name, b)
        assert np.array_equal(grad, g["dphi"][b]), (name, b)


def _interior(x):
    return np.all((x > 0.2) * (x < 0.8), axis=0)


@pytest.mark.parametrize("name,M,E,normal", [
    ("facet_oriented_tri", "MeshTri", "ElementTriP2", [1., 0.2]),
    ("facet_oriented_tet", "MeshTet", "ElementTetP1", [1., 0.2, -0.1])])
def test_oriented_facet_sets_match_reference(name, M, E, normal):
    """OrientedBoundary (Mesh.facets_around, facets_satisfying(normal=...)) through FacetBasis
    (facet_basis.py:84-89): the facet sets, their orientation, the elements the traces and the
    normals are taken from, and - through the shipped geometry kernel on the host - normals
    and dx, all equal to the reference's."""
    g = load(name)
    m = getattr(fem, M)(g["p"], g["t"])
    inside = m.elements_satisfying(lambda x: _interior(x) * (x[0] < 0.55))
    assert np.array_equal(inside, g["inside"])
    sets = {"around": m.facets_around(inside), "around_flip": m.facets_around(inside, flip=True),
            "normal": m.facets_satisfying(lambda x: _interior(x) * (x[0] > 0.3) * (x[0] < 0.7),
                                          normal=np.array(normal))}
    lib = host_facet.lib()
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)          # noqa: E731
    dim = m.dim()
    p, t, facets = np.ascontiguousarray(m.p), i32(m.t), i32(m.facets)
    for key, ob in sets.items():
        assert isinstance(ob, fem.OrientedBoundary)
        assert np.array_equal(np.asarray(ob), g[key + "_find"])
        assert np.array_equal(ob.ori, g[key + "_ori"])
        for side in (0, 1):
            k = "{}_s{}".format(key, side)
            fb = fem.FacetBasis(m, getattr(fem, E)(), facets=ob, side=side)
            assert np.array_equal(fb.tind, g[k + "_tind"])
            assert np.array_equal(fb.tind_normals, g[k + "_tind_normals"])
            assert fb.with_element(getattr(fem, E)()).tind.tolist() == fb.tind.tolist()
            nf, nqp = fb.nelems, fb.nqp
            find, tind, tind_n = i32(fb.find), i32(fb.tind), i32(fb.tind_normals)
            lfacet = i32(np.argmax(m.t2f[:, fb.tind_normals] == np.asarray(fb.find)[None, :],
                                   axis=0))
            Xb, Wb = np.ascontiguousarray(fb.X), np.ascontiguousarray(fb.W)
            sp = _lib.SkbSpace()
            sp.dim, sp.nnodes, sp.mapping = dim, m.t.shape[0], _lib.SKB_MAP_AFFINE
            sp.nbs, sp.ncomp, sp.nqp = fb.nbs, fb.ncomp, nqp
            sp.npts, sp.nel_total, sp.nel = m.p.shape[1], m.t.shape[1], nf
            sp.p, sp.t, sp.tind = p.ctypes.data, t.ctypes.data, tind.ctypes.data
            x, Y, nrm = (np.full((dim, nf, nqp), np.nan) for _ in range(3))
            dx, detabs = np.full((nf, nqp), np.nan), np.full((nf, nqp), np.nan)
            lib.host_facet_geometry(C.byref(sp), _p(facets), C.c_int64(facets.shape[1]), _p(find),
                                    _p(tind), _p(tind_n), _p(lfacet), C.c_int64(nf), _p(Xb),
                                    _p(Wb), C.c_int(nqp), _p(x), _p(Y), _p(dx), _p(nrm),
                                    _p(detabs))
            assert np.array_equal(nrm, g[k + "_normals"]), k
            assert np.array_equal(dx, g[k + "_dx"]), k
            # divergence theorem on the closed oriented surfaces: dim * volume, sign by side of n
            if key != "normal":
                div = float(np.sum(np.sum(nrm * x, axis=0) * dx))
                np.testing.assert_allclose(div, float(g[k + "_divthm"]), rtol=1e-12)
