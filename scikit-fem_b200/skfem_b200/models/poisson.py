"""Poisson-type library forms - the integrands of skfem/models/poisson.py:7-24.

Every entry of the table below becomes a form object that carries a ``native`` tag
``(kind, kernel id, parameters, field type)``: ``Form.assemble`` recognises the tag and
launches the CUDA kernel written for that integrand (geometry, push-forward and
quadrature reduction in one pass) instead of tracing the Python callable.  The callable
stays a valid definition - it is what the traced path executes where no dedicated kernel
applies (two different bases, a FacetBasis, extra parameters).
"""
from .. import _lib
from ..form import BilinearForm, LinearForm
from ..helpers import ddot, dot, grad


def _library_form(wrapper, name, integrand, native):
    integrand.__name__ = integrand.__qualname__ = name
    form = wrapper(integrand)
    form.native = native
    return form


laplace = _library_form(
    BilinearForm, "laplace", lambda u, v, w: dot(grad(u), grad(v)),
    ("bilinear", _lib.FORM_LAPLACE, None, "scalar"))

vector_laplace = _library_form(
    BilinearForm, "vector_laplace", lambda u, v, w: ddot(grad(u), grad(v)),
    ("bilinear", _lib.FORM_VECTOR_LAPLACE, None, "vector"))

mass = _library_form(
    BilinearForm, "mass", lambda u, v, w: u * v,
    ("bilinear", _lib.FORM_MASS, None, "scalar"))

unit_load = _library_form(
    LinearForm, "unit_load", lambda v, w: v,
    ("linear", _lib.LFORM_UNIT_LOAD, None, "scalar"))
