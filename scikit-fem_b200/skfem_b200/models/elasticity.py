"""Linear elasticity forms (skfem/models/elasticity.py:7-53).

The material helpers evaluate the same floating-point expressions in the same order as
the reference, so parameters derived through them are bit-identical; the weak form
carries a ``native`` tag so ``Form.assemble`` launches the dedicated CUDA kernel.
"""
from .. import _lib, helpers as H
from ..form import BilinearForm


def lame_parameters(E, nu):
    """(Young's modulus, Poisson ratio) -> (first Lame parameter, shear modulus)."""
    one_plus = 1. + nu
    lam = E * nu / (one_plus * (1. - 2. * nu))
    mu = E / (2. * one_plus)
    return lam, mu


def plane_stress(E, nu):
    """Effective (E, nu) that turn the plane-strain formulas into plane stress."""
    one_plus = 1. + nu
    return E * (1. + 2. * nu) / one_plus ** 2, nu / one_plus


def linear_stress(Lambda=1., Mu=1.):
    """Isotropic Hooke law: strain tensor field -> stress tensor field."""
    return lambda T: 2. * Mu * T + Lambda * H.eye(H.trace(T), T.shape[0])


def linear_elasticity(Lambda=1., Mu=1.):
    """a(u, v) = (C eps(u), eps(v)).  The kernel receives Lambda and the host-evaluated
    product 2.*Mu - the same two scalars numpy broadcasts in ``linear_stress``."""
    stress = linear_stress(Lambda, Mu)

    def weakform(u, v, w):
        return H.ddot(stress(H.sym_grad(u)), H.sym_grad(v))

    form = BilinearForm(weakform)
    form.native = ("bilinear", _lib.FORM_ELASTICITY,
                   (float(Lambda), 2. * float(Mu)), "vector")
    return form
