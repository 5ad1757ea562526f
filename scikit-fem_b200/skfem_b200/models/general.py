"""Integrands "too general for a specific model" (the four forms of
skfem/models/general.py:7-27).  No dedicated kernels: they always take the traced path,
typically with two different bases (velocity / pressure) or a field passed as ``w['w']``.
"""
from ..form import BilinearForm, LinearForm
from ..helpers import curl, div, dot


def _traced(wrapper, name, integrand):
    integrand.__name__ = integrand.__qualname__ = name
    return wrapper(integrand)


# b(u, p) = (div u, p): the divergence constraint of Stokes-type problems
divu = divergence = _traced(BilinearForm, "divu", lambda u, v, w: div(u) * v)
# (curl u, v)
curluv = _traced(BilinearForm, "curluv", lambda u, v, w: dot(curl(u), v))
# right-hand sides with a given field w['w']: (curl v, w) and (v, curl w)
rot = _traced(LinearForm, "rot", lambda v, w: dot(curl(v), w['w']))
vrot = _traced(LinearForm, "vrot", lambda v, w: dot(v, curl(w['w'])))
