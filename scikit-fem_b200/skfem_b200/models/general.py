"""Forms too general to put into a specific model (skfem/models/general.py:1-27);
they run through the traced path."""
from ..form import BilinearForm, LinearForm
from ..helpers import curl, div, dot


@BilinearForm
def divu(u, v, w):
    return div(u) * v


divergence = divu


@BilinearForm
def curluv(u, v, w):
    return dot(curl(u), v)


@LinearForm
def rot(v, w):
    return dot(curl(v), w['w'])


@LinearForm
def vrot(v, w):
    return dot(v, curl(w['w']))
