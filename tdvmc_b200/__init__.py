"""tdvmc_b200 -- B200-native walker-ensemble hot path for time-dependent variational Monte Carlo.

Product path: ``capi`` (ctypes over libtdvmc_b200.so, hand-written sm_100a CUDA) and ``ensemble``
(the reference's per-rank estimator loops on top of it).  ``systems``/``splines`` build the
system description the C ABI takes.  Nothing in this package imports or falls back to ``oracle/``.
"""
from . import estimators, splines, systems  # noqa: F401

__all__ = ["capi", "ensemble", "estimators", "splines", "systems"]
