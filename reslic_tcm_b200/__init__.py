"""reslic_tcm_b200 — B200-native (sm_100a) entropy-model hot path for ResLIC_TCM.

Only what the path needs: ``csrc/`` (CUDA kernels + C ABI), the ctypes binding, and the
host-side mirror of the reference's entropy-model operator interface.
"""
from .entropy_models import EntropyBottleneck, EntropyModel, GaussianConditional, LowerBound  # noqa: F401
from ._cabi import ReslicError  # noqa: F401

__version__ = "0.1.0"
