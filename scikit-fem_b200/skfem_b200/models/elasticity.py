"""Linear elasticity forms (skfem/models/elasticity.py:7-53)."""
from ..form import BilinearForm
from ..helpers import ddot, trace, sym_grad, eye
from .._lib import FORM_ELASTICITY


def lame_parameters(E, nu):
    """Young's modulus, Poisson ratio -> (lambda, mu)."""
    return (E * nu / ((1. + nu) * (1. - 2. * nu)), E / (2. * (1. + nu)))


def plane_stress(E, nu):
    return (E * (1. + 2. * nu) / (1. + nu) ** 2, nu / (1. + nu))


def linear_stress(Lambda=1., Mu=1.):
    def C(T):
        return 2. * Mu * T + Lambda * eye(trace(T), T.shape[0])
    return C


def linear_elasticity(Lambda=1., Mu=1.):
    C = linear_stress(Lambda, Mu)

    @BilinearForm
    def weakform(u, v, w):
        return ddot(C(sym_grad(u)), sym_grad(v))

    # the kernel receives Lambda and the host-evaluated product 2.*Mu, the
    # same two scalars numpy broadcasts (elasticity.py:40)
    weakform.native = ("bilinear", FORM_ELASTICITY, (float(Lambda), 2. * float(Mu)), "vector")
    return weakform
