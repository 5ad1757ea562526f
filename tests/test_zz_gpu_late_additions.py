"""GPU tests added after the last GPU session of round 2 (the round's GPU budget was spent).

What they exercise on the device is verified on the CPU from the shipped sources - the Hex2
kernels in tests/test_hex_kernels_cpu.py (scalar kernel bit for bit, sum-factorised kernel rtol
1e-12 on this very fixture), the facet kernels with oriented facet sets in
tests/test_facet_kernels_cpu.py - but these device runs themselves have not been observed yet,
so the file sorts last: everything that was observed green on a B200 runs before it."""
import numpy as np
import pytest

import skfem_b200 as fem
from cases import load
from product import forms, mesh_from

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.mark.parametrize("sumfact", [True, False])
def test_hex2_value_parity_on_anisotropic_boxes(sumfact):
    """tests/golden/hex2_boxes3.npz (27 boxes with power-of-two edge ratios, real reference):
    pattern bit-exact, values rtol 1e-12 for both Hex2 kernels; the scalar kernel bit-exact."""
    from skfem_b200 import _lib, form as F
    g = load("hex2_boxes3")
    b = fem.Basis(mesh_from(g, "hex"), fem.ElementHex2())
    assert np.array_equal(b.element_dofs, g["element_dofs"])
    fs = forms(False)
    F.set_options(hex_sumfact=sumfact)
    try:
        for f in ["laplace", "mass"]:
            A = fs[f].assemble(b)
            assert np.array_equal(A.indptr, g[f + "_indptr"])
            assert np.array_equal(A.indices, g[f + "_indices"])
            ref = g[f + "_data"]
            np.testing.assert_allclose(A.data, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())
            loc = g[f + "_local"]
            got = fs[f].elemental(b).data
            np.testing.assert_allclose(got, loc, rtol=RTOL, atol=RTOL * np.abs(loc).max())
            try:
                _lib.lib().skb_debug_flags(8)           # scalar kernel, reference order
                assert np.array_equal(fs[f].elemental(b).data, loc), f
            finally:
                _lib.lib().skb_debug_flags(0)
    finally:
        F.set_options(hex_sumfact=True)


def _elem(fem, ename):
    return {"tri_p2": fem.ElementTriP2, "tet_p1": fem.ElementTetP1}[ename]()


@pytest.mark.parametrize("name,refdom,ename,normal", [
    ("facet_oriented_tri", "tri", "tri_p2", [1., 0.2]),
    ("facet_oriented_tet", "tet", "tet_p1", [1., 0.2, -0.1])])
def test_gpu_facet_oriented_sets(name, refdom, ename, normal):
    """FacetBasis on OrientedBoundary sets (Mesh.facets_around, facets_satisfying(normal=));
    golden vectors by the real reference (tools/gen_golden_oriented.py)."""
    from skfem_b200.helpers import dot, grad
    g = load(name)
    m = (fem.MeshTri if refdom == "tri" else fem.MeshTet)(g["p"], g["t"])
    e = _elem(fem, ename)

    def interior(x):
        return np.all((x > 0.2) * (x < 0.8), axis=0)
    inside = m.elements_satisfying(lambda x: interior(x) * (x[0] < 0.55))
    sets = {"around": m.facets_around(inside), "around_flip": m.facets_around(inside, flip=True),
            "normal": m.facets_satisfying(lambda x: interior(x) * (x[0] > 0.3) * (x[0] < 0.7),
                                          normal=np.array(normal))}
    flow = fem.BilinearForm(lambda u, v, w: dot(grad(u), w.n) * v + u * v)
    divthm = fem.Functional(lambda w: dot(w.n, w.x))
    for key, ob in sets.items():
        assert np.array_equal(np.asarray(ob), g[key + "_find"])
        assert np.array_equal(ob.ori, g[key + "_ori"])
        for side in (0, 1):
            k = "{}_s{}".format(key, side)
            fb = fem.FacetBasis(m, e, facets=ob, side=side)
            assert np.array_equal(fb.tind, g[k + "_tind"])
            assert np.array_equal(fb.normals.numpy(), g[k + "_normals"])
            assert np.array_equal(fb.dx, g[k + "_dx"])
            A = flow.assemble(fb)
            assert np.array_equal(A.indptr, g[k + "_indptr"])
            assert np.array_equal(A.indices, g[k + "_indices"])
            ref = g[k + "_data"]
            np.testing.assert_allclose(A.data, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
            np.testing.assert_allclose(divthm.assemble(fb), float(g[k + "_divthm"]), rtol=1e-12)


