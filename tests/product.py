"""Rebuild golden cases with the product (skfem_b200)."""
import numpy as np

import skfem_b200 as fem
from skfem_b200.helpers import dot, grad
from skfem_b200.models.poisson import laplace, mass, unit_load, vector_laplace
from skfem_b200.models.elasticity import linear_elasticity

from cases import LAME

MESH = {"tri": fem.MeshTri, "tet": fem.MeshTet, "hex": fem.MeshHex}
ELEM = {"tri_p1": fem.ElementTriP1, "tri_p2": fem.ElementTriP2, "tet_p1": fem.ElementTetP1,
        "tet_p2": fem.ElementTetP2,
        "hex1": fem.ElementHex1, "hex2": fem.ElementHex2}


@fem.BilinearForm
def user_aniso(u, v, w):
    return (1. + w.x[0] * w.x[1]) * dot(grad(u), grad(v)) + 3. * u * v


@fem.LinearForm
def user_load(v, w):
    return np.sin(3. * w.x[0]) * v + w.x[1] * v


vector_mass = fem.BilinearForm(lambda u, v, w: dot(u, v))


def forms(vector):
    return dict(laplace=laplace, mass=vector_mass if vector else mass,
                vector_laplace=vector_laplace, elasticity=linear_elasticity(*LAME),
                user_aniso=user_aniso, unit_load=unit_load, user_load=user_load)


def mesh_from(g, refdom):
    m = MESH[refdom](g["p"], g["t"])
    if refdom == "tri":
        assert np.array_equal(m.t, g["t"])  # already column-sorted
    return m


def element_from(name, vector):
    e = ELEM[name]()
    return fem.ElementVector(e) if vector else e
