"""pygam_b200 -- a B200-native PIRLS fitting engine behind the pyGAM model API.

Same public names as ``pygam`` (``/root/reference/pygam/__init__.py:5-29``); the hot path runs
in hand-written sm_100a CUDA kernels (``pygam_b200/csrc``) reached through the C ABI of
``include/pgb.h``.  There is no CPU fallback.
"""

from .engine import DeviceData
from .gam import (GAM, ExpectileGAM, GammaGAM, InvGaussGAM, LinearGAM, LogisticGAM, NotPositiveDefiniteError,
                  OptimizationError, PoissonGAM)
from .terms import f, intercept, l, s, te

__all__ = ["GAM", "LinearGAM", "LogisticGAM", "GammaGAM", "PoissonGAM", "InvGaussGAM", "ExpectileGAM",
           "l", "s", "f", "te", "intercept", "DeviceData", "NotPositiveDefiniteError", "OptimizationError"]

__version__ = "0.1.0"
