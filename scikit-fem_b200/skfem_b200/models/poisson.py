"""Poisson-type forms (skfem/models/poisson.py:7-24).

Each library form carries a ``native`` tag: ``Form.assemble`` recognises it and
dispatches to the fused CUDA kernel for that integrand (geometry, push-forward
and quadrature reduction in one pass) instead of tracing the Python callable.
The callables remain valid definitions and are what the traced path executes
when a fused kernel does not apply (e.g. two different bases).
"""
from ..form import BilinearForm, LinearForm
from ..helpers import grad, dot, ddot
from .._lib import FORM_LAPLACE, FORM_MASS, FORM_VECTOR_LAPLACE, LFORM_UNIT_LOAD


@BilinearForm
def laplace(u, v, _):
    return dot(grad(u), grad(v))


@BilinearForm
def vector_laplace(u, v, _):
    return ddot(grad(u), grad(v))


@BilinearForm
def mass(u, v, _):
    return u * v


@LinearForm
def unit_load(v, _):
    return v


laplace.native = ("bilinear", FORM_LAPLACE, None, "scalar")
vector_laplace.native = ("bilinear", FORM_VECTOR_LAPLACE, None, "vector")
mass.native = ("bilinear", FORM_MASS, None, "scalar")
unit_load.native = ("linear", LFORM_UNIT_LOAD, None, "scalar")
