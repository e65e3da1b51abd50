"""bayesrul_b200 -- B200-native (sm_100a) implementation of bayesrul's Monte-Carlo variational-BNN hot path.

Public surface:
  Engine / Noise      tensor-level host of the C ABI (include/bayesrul_b200.h)
  build()             in-tree nvcc build of lib/libbayesrul_b200.so
The reference-facing mirrors (BNN, HNN, tyxe / pyro shims) live in bayesrul_b200.compat.
"""
from .build import build  # noqa: F401
from .engine import Engine, Noise, net_info  # noqa: F401

__all__ = ["Engine", "Noise", "net_info", "build"]
