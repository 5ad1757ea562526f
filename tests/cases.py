"""Golden cases (tests/golden/*.npz made by tools/gen_golden.py from the real
reference) and how to rebuild each of them with the oracle / the product."""
import os
from types import SimpleNamespace

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fixture -> (refdom, element, vector?, bilinear forms, linear forms, has_local)
CASES = {
    "c1_tri_p1_refined4": ("tri", "tri_p1", False,
                           ["laplace", "mass", "user_aniso"],
                           ["unit_load", "user_load"], True),
    "tri_p1_two_triangles": ("tri", "tri_p1", False, ["laplace", "mass"],
                             ["unit_load"], True),
    "tri_p2_morphed3": ("tri", "tri_p2", False, ["laplace", "mass", "user_aniso"],
                        ["unit_load", "user_load"], True),
    "tet_p1_tensor6": ("tet", "tet_p1", False,
                       ["laplace", "mass", "user_aniso"],
                       ["unit_load", "user_load"], True),
    "tet_p1_tensor_nonuniform": ("tet", "tet_p1", False, ["laplace", "mass"],
                                 ["unit_load"], True),
    "tet_p1_ball2": ("tet", "tet_p1", False, ["laplace", "mass"],
                     ["unit_load"], True),
    "tet_p1_refined3": ("tet", "tet_p1", False, ["laplace"], ["unit_load"],
                        True),
    "tet_p1_morphed5": ("tet", "tet_p1", False, ["laplace", "mass"],
                        ["unit_load"], True),
    "tet_p2_tensor3": ("tet", "tet_p2", False, ["laplace", "mass"],
                       ["unit_load"], True),
    "tet_p2_morphed3": ("tet", "tet_p2", False,
                        ["laplace", "mass", "user_aniso"],
                        ["unit_load", "user_load"], True),
    "tet_vp2_elasticity2": ("tet", "tet_p2", True,
                            ["elasticity", "vector_laplace"], [], True),
    "tet_vp2_elasticity_morphed2": ("tet", "tet_p2", True, ["elasticity"], [],
                                    True),
    "tet_vp1_elasticity4": ("tet", "tet_p1", True,
                            ["elasticity", "vector_laplace", "mass"], [],
                            True),
    "hex1_tensor3": ("hex", "hex1", False, ["laplace", "mass"], ["unit_load"],
                     True),
    "hex1_morphed3": ("hex", "hex1", False, ["laplace", "mass"],
                      ["unit_load"], True),
}

LAME = (1e3 * 0.3 / ((1. + 0.3) * (1. - 2. * 0.3)), 1e3 / (2. * (1. + 0.3)))


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def mesh_of(g, refdom):
    return SimpleNamespace(p=np.ascontiguousarray(g["p"]),
                           t=np.ascontiguousarray(g["t"]), refdom=refdom)
