"""sde4mbrl_b200 -- B200-native (sm_100a) hot path of wuwushrek/sde4mbrl.

Batched Monte-Carlo Euler-Maruyama particle rollouts of the controlled neural SDE and the fused
reverse-mode pass the APG-MPC solver differentiates through, behind the reference's own
``nsde.py`` sampling API and ``apg.py`` solver entry point.  All arithmetic of the path runs in
hand-written CUDA kernels (``libsde4b200.so``, C-ABI in ``include/sde4b200.h``); there is no CPU
fallback.
"""
from . import _lib
from ._lib import Sde4Error

__all__ = ["_lib", "Sde4Error", "nsde", "apg", "random", "costs", "sde_models", "engine"]
__version__ = "0.1.0"
