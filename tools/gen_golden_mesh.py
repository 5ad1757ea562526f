"""Golden vectors for the mesh constructors that have no other fixture: uniform refinement of
tetrahedral and hexahedral meshes (mesh_tet_1.py:84-126, mesh_hex_1.py:57-95), produced by the
real reference.  Run in the build container only:

    PYTHONPATH=/root/reference python tools/gen_golden_mesh.py
"""
import os

import numpy as np
import skfem

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "mesh_refined.npz")


def main():
    x = np.array([0., 0.2, 0.55, 1.])
    y = np.array([0., 0.4, 1.])
    z = np.array([0., 0.1, 0.35, 0.8, 1.])
    out = {"x": x, "y": y, "z": z}
    cases = {
        "tet_default_r2": skfem.MeshTet().refined(2),
        "tet_tensor_r1": skfem.MeshTet.init_tensor(x, y, z).refined(1),
        "hex_default_r2": skfem.MeshHex().refined(2),
        "hex_tensor_r1": skfem.MeshHex.init_tensor(x, y, z).refined(1),
    }
    for name, m in cases.items():
        out[name + "_p"] = m.p
        out[name + "_t"] = m.t
    # selections and mesh parameters (mesh.py:251-275,402-424,476-493, mesh_2d/3d.py params)
    sel = {"tet": skfem.MeshTet.init_tensor(x, y, z), "tri": skfem.MeshTri().refined(3),
           "hex": skfem.MeshHex.init_tensor(x, y, z)}
    for name, m in sel.items():
        ms = m.with_subdomains({"low": lambda q: q[0] < 0.5, "pick": np.array([0, 2])})
        out[name + "_params"] = m.params()
        out[name + "_interior_nodes"] = m.interior_nodes()
        out[name + "_nodes_low"] = m.nodes_satisfying(lambda q: q[0] < 0.5)
        out[name + "_bnodes_low"] = m.nodes_satisfying(lambda q: q[0] < 0.5, boundaries_only=True)
        out[name + "_elements_low"] = m.elements_satisfying(lambda q: q[0] < 0.5)
        out[name + "_norm_names"] = ms.normalize_elements(["low", "pick"])
        out[name + "_norm_list"] = ms.normalize_elements([4, 1, 1])
        basis = skfem.Basis(ms, {"tet": skfem.ElementTetP2, "tri": skfem.ElementTriP2,
                                 "hex": skfem.ElementHex2}[name](), elements="low")
        out[name + "_sub_tind"] = basis.tind
        out[name + "_sub_nelems"] = basis.nelems
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
