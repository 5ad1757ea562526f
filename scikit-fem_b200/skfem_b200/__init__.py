"""skfem_b200 - B200-native finite element assembly behind the scikit-fem API.

    from skfem_b200 import *
    from skfem_b200.models.poisson import laplace, unit_load
    m = MeshTet.init_tensor(x, y, z)
    basis = Basis(m, ElementTetP1())
    A = laplace.assemble(basis)          # scipy.sparse.csr_matrix
    b = unit_load.assemble(basis)        # numpy.ndarray

Host side: thin Python; device side: hand-written sm_100a CUDA kernels behind
the C ABI in include/skfem_b200.h.  There is no CPU fallback.
"""
from .mesh import (Mesh, MeshTri, MeshTet, MeshHex, MeshTri1, MeshTet1, MeshHex1,
                   OrientedBoundary)
from .element import (Element, ElementH1, ElementTriP1, ElementTriP2, ElementTetP1,
                      ElementTetP2, ElementHex1, ElementHex2, ElementVector)
from .mapping import MappingAffine, MappingIsoparametric
from .dofs import Dofs, DofsView
from .quadrature import get_quadrature
from .basis import AbstractBasis, CellBasis, Basis
from .facet_basis import FacetBasis, BoundaryFacetBasis, InteriorFacetBasis
from .field import DiscreteField, DeviceArray, asdevice
from .form import (Form, BilinearForm, LinearForm, Functional, COOData, DeviceCSR,
                   FormExtraParams, asm)
from .utils import enforce, condense, solve, solver_iter_pcg
from . import helpers, models, quadrature, utils

InteriorBasis = CellBasis  # deprecated aliases kept by the reference
ExteriorFacetBasis = FacetBasis

__version__ = "0.1.0"

__all__ = [
    "Mesh", "MeshTri", "MeshTet", "MeshHex", "MeshTri1", "MeshTet1", "MeshHex1",
    "Element", "ElementH1", "ElementTriP1", "ElementTriP2", "ElementTetP1", "ElementTetP2",
    "ElementHex1", "ElementHex2", "ElementVector", "MappingAffine", "MappingIsoparametric",
    "Dofs", "DofsView", "get_quadrature", "AbstractBasis", "CellBasis", "Basis", "InteriorBasis",
    "FacetBasis", "BoundaryFacetBasis", "ExteriorFacetBasis", "InteriorFacetBasis",
    "DiscreteField", "DeviceArray", "asdevice", "Form", "BilinearForm", "LinearForm",
    "Functional", "COOData", "DeviceCSR", "FormExtraParams", "asm", "helpers", "models",
    "enforce", "condense", "solve", "solver_iter_pcg", "utils", "OrientedBoundary",
]
