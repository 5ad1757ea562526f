"""Form library (the reference's ``skfem.models``)."""
from . import poisson, elasticity  # noqa: F401
