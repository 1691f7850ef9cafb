"""pmf -> quantised CDF rows for the rANS tables (host C++; setup code, not per-element).

Replaces ``compressai._CXX.pmf_to_quantized_cdf`` as wrapped by the reference at
``src/entropy_models/coder.py:53-56``.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from . import _cabi


def pmf_to_quantized_cdf(pmf: Tensor, precision: int = 16) -> Tensor:
    lib = _cabi.load()
    p = pmf.detach().to(device="cpu", dtype=torch.float32).contiguous().reshape(-1)
    out = torch.empty(p.numel() + 1, dtype=torch.int32)
    code = lib.reslic_pmf_to_quantized_cdf(p.data_ptr(), p.numel(), int(precision), out.data_ptr())
    if code != 0:
        raise ValueError(lib.reslic_last_error().decode())
    return out
