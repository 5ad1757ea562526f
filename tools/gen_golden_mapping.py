"""Golden vectors for the Element / Mapping contract: Mapping.F / DF / invDF / detDF
(signed) / invF and Element.gbasis(mapping, X, i, tind), produced by the REAL reference
(scikit-fem 12.0.1, /root/reference).

    python tools/gen_golden_mapping.py        -> tests/golden/mapping_contract.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import skfem as fem  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
out = {}

# affine: morphed tets, one of them inverted (negative determinant)
x = np.linspace(0, 1, 4)
m = fem.MeshTet.init_tensor(x, np.linspace(0, 1, 3), x)
q = m.p.copy()
q[0] = m.p[0] + 0.03 * np.sin(7 * m.p[1])
q[1] = m.p[1] + 0.02 * m.p[2] ** 2
t = m.t.copy()
t[[1, 2], 5] = t[[2, 1], 5]                     # swap two vertices: det < 0 for element 5
m = fem.MeshTet(q, t)
X = np.array([[0.1, 0.3, 0.25], [0.2, 0.3, 0.25], [0.3, 0.1, 0.25]])
tind = np.array([0, 5, 7, 20], dtype=np.int32)
mp = m._mapping()
out.update(tet_p=m.p, tet_t=m.t, tet_X=X, tet_tind=tind,
           tet_F=mp.F(X), tet_DF=mp.DF(X), tet_invDF=mp.invDF(X), tet_detDF=mp.detDF(X),
           tet_detDF_tind=mp.detDF(X, tind), tet_invDF_tind=mp.invDF(X, tind))
xg = mp.F(X)
out["tet_invF"] = mp.invF(xg)
e = fem.ElementTetP2()
for i in (0, 4, 9):
    f = e.gbasis(mp, X, i, tind)[0]
    out["tet_p2_gb{}_value".format(i)] = np.array(f.value)
    out["tet_p2_gb{}_grad".format(i)] = np.array(f.grad)
ev = fem.ElementVector(fem.ElementTetP1())
f = ev.gbasis(mp, X, 7)[0]
out["tet_vp1_gb7_value"] = np.array(f.value)
out["tet_vp1_gb7_grad"] = np.array(f.grad)

# isoparametric: morphed hexahedra
xh = np.linspace(0, 1, 3)
mh = fem.MeshHex.init_tensor(xh, xh, xh)
qh = mh.p.copy()
qh[0] = mh.p[0] + 0.05 * np.sin(3 * mh.p[1]) * mh.p[2]
qh[2] = mh.p[2] + 0.04 * mh.p[0] * mh.p[1]
mh = fem.MeshHex(qh, mh.t)
Xh = np.array([[0.2, 0.7, 0.5], [0.3, 0.6, 0.5], [0.9, 0.1, 0.5]])
mph = mh._mapping()
out.update(hex_p=mh.p, hex_t=mh.t, hex_X=Xh, hex_F=mph.F(Xh), hex_DF=mph.DF(Xh),
           hex_invDF=mph.invDF(Xh), hex_detDF=mph.detDF(Xh))
fh = fem.ElementHex1().gbasis(mph, Xh, 6)[0]
out["hex1_gb6_value"] = np.array(fh.value)
out["hex1_gb6_grad"] = np.array(fh.grad)
np.savez_compressed(os.path.join(OUT, "mapping_contract.npz"), **out)
print({k: v.shape for k, v in out.items()})
print("negative determinants:", (out["tet_detDF"] < 0).sum())
