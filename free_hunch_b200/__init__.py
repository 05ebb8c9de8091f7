"""free_hunch_b200: B200-native (sm_100a) implementation of the Free Hunch guided-sampling hot path.

Public surface mirrors the reference modules it replaces:
  conditioning_mechanisms.choose_conditioning_mechanism('free_hunch')   (conditioning_utils/conditioning_mechanisms.py)
  covariance.CovarianceHessianBFGS / CovarianceHessianBFGSDCT           (conditioning_utils/online_update_bfgs.py)
  measurements.get_operator(...)                                        (measurement_utils/measurements.py)
  engine.cg_solve / conditioning_mechanisms.choose_solver               (conditioning_utils/cg.py + mat solvers)
The compute lives in libfreehunch_b200.so (C ABI in include/free_hunch_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (import fails loudly later, on first use, if the library is missing)

__all__ = ["engine", "covariance", "measurements", "conditioning_mechanisms", "sampler"]
