"""Form library (the reference's ``skfem.models``)."""
from . import poisson, elasticity, general  # noqa: F401
