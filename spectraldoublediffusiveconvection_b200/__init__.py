"""B200-native ensemble time-stepper for spherical-shell double-diffusive convection (hot path only).

Importing the package does not touch the GPU; creating an EnsemblePlan (or calling any operator of the
Matrix_Operators / Transforms drop-in modules) loads libsddc_b200.so and fails loudly if it is missing.
"""
from .operators import RadialOperators, cheb_radial  # noqa: F401

__all__ = ["RadialOperators", "cheb_radial", "EnsemblePlan", "transform"]


def __getattr__(name):
    if name in ("EnsemblePlan", "transform", "SddcError"):
        from . import plan
        return getattr(plan, name)
    raise AttributeError(name)
