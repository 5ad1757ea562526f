"""Element / Mapping contract of the drop-in boundary: Mapping.F / DF / invDF /
detDF (signed) / invF and Element.gbasis(mapping, X, i, tind) against the real reference
(tests/golden/mapping_contract.npz, tools/gen_golden_mapping.py) - bit for bit."""
import numpy as np
import pytest

from cases import load

pytestmark = pytest.mark.gpu


def test_mapping_arrays_match_the_reference():
    import skfem_b200 as fem
    g = load("mapping_contract")
    m = fem.MeshTet(g["tet_p"], g["tet_t"])
    mp = m._mapping()
    X, tind = g["tet_X"], g["tet_tind"]
    assert np.array_equal(mp.F(X), g["tet_F"])
    assert np.array_equal(mp.DF(X), g["tet_DF"])
    assert np.array_equal(mp.invDF(X), g["tet_invDF"])
    det = mp.detDF(X)
    assert np.array_equal(det, g["tet_detDF"]) and (det < 0).any()      # signed
    assert np.array_equal(mp.detDF(X, tind), g["tet_detDF_tind"])
    assert np.array_equal(mp.invDF(X, tind), g["tet_invDF_tind"])
    assert np.array_equal(mp.invF(g["tet_F"]), g["tet_invF"])
    mh = fem.MeshHex(g["hex_p"], g["hex_t"])
    mph = mh._mapping()
    Xh = g["hex_X"]
    assert np.array_equal(mph.F(Xh), g["hex_F"])
    assert np.array_equal(mph.DF(Xh), g["hex_DF"])
    assert np.array_equal(mph.invDF(Xh), g["hex_invDF"])
    assert np.array_equal(mph.detDF(Xh), g["hex_detDF"])
    flat = fem.MeshHex(0.0 * g["hex_p"], g["hex_t"])     # (the mapping holds its mesh weakly)
    with pytest.raises(Exception, match="Zero Jacobian determinant"):
        flat._mapping().detDF(Xh)


def test_element_gbasis_matches_the_reference():
    import skfem_b200 as fem
    g = load("mapping_contract")
    m = fem.MeshTet(g["tet_p"], g["tet_t"])
    mp = m._mapping()
    X, tind = g["tet_X"], g["tet_tind"]
    e = fem.ElementTetP2()
    for i in (0, 4, 9):
        (f,) = e.gbasis(mp, X, i, tind)
        assert np.array_equal(np.asarray(f.numpy()), g["tet_p2_gb{}_value".format(i)])
        assert np.array_equal(np.asarray(f.grad.numpy()), g["tet_p2_gb{}_grad".format(i)])
    (f,) = fem.ElementVector(fem.ElementTetP1()).gbasis(mp, X, 7)
    assert np.array_equal(np.asarray(f.numpy()), g["tet_vp1_gb7_value"])
    assert np.array_equal(np.asarray(f.grad.numpy()), g["tet_vp1_gb7_grad"])
    mh = fem.MeshHex(g["hex_p"], g["hex_t"])
    (fh,) = fem.ElementHex1().gbasis(mh._mapping(), g["hex_X"], 6)
    assert np.array_equal(np.asarray(fh.numpy()), g["hex1_gb6_value"])
    assert np.array_equal(np.asarray(fh.grad.numpy()), g["hex1_gb6_grad"])
    with pytest.raises(ValueError):
        e.gbasis(mp, X, 10)
