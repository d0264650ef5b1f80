"""rlshaders_b200 -- B200-native (sm_100a) batched BSDF hot path of shihchinw/rlShaders.

The package is a thin host-side mirror of the reference's sampler interface over the
C ABI in include/rls_b200.h (librls_b200.so, hand-written CUDA).  There is no CPU
fallback: creating a Context without the compiled library or without a B200-class
device raises.
"""
from . import _abi  # noqa: F401  (pure ctypes layout, safe to import anywhere)

__all__ = ["_abi"]
