"""Host-side mirror of the reference's ``RateDistortionLoss`` (src/training/loss.py:6-35) over the fused rate
reduction: same constructor, same ``forward(output, target, p=None)`` and the same keys in the result.

The rate term ``sum(log(L).sum() / (-ln 2 * num_pixels))`` (loss.py:22-25) is one read pass per likelihood tensor
(``ops.rate_from_likelihood``: deterministic per-image sums on the GPU) instead of a log tensor plus a reduction,
and its backward is the single elementwise ``g / L``.  The distortion term stays ``nn.MSELoss`` (already one fused
reduction in torch).  ``type='ms_ssim'`` needs the third-party ``pytorch_msssim`` of ``utils/helper.compute_msssim``
and is not part of this path: it raises.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops

__all__ = ["RateDistortionLoss", "rate_bits"]


class _RateBits(torch.autograd.Function):
    """bits[b] = -sum log2 L[b]  (float64 [B]);  d bits[b] / d L = -1 / (ln 2 * L)."""

    @staticmethod
    def forward(ctx, lik: Tensor) -> Tensor:
        ctx.save_for_backward(lik)
        return ops.rate_from_likelihood(lik.detach())

    @staticmethod
    def backward(ctx, g_bits: Tensor):
        (lik,) = ctx.saved_tensors
        coef = (g_bits.double() * (-1.0 / math.log(2.0))).to(lik.dtype)
        return coef.reshape(-1, *([1] * (lik.dim() - 1))) / lik


def rate_bits(likelihoods: Tensor) -> Tensor:
    """Per-image bits of a likelihood tensor [B, ...], differentiable w.r.t. the likelihoods."""
    return _RateBits.apply(likelihoods)


class RateDistortionLoss(nn.Module):
    """Custom rate distortion loss with a Lagrangian parameter (loss.py:6-35)."""

    def __init__(self, lmbda=[1e-2], type="mse"):
        super().__init__()
        self.mse = nn.MSELoss()
        self.lmbda = lmbda
        self.type = type

    def forward(self, output: Dict, target: Tensor, p: Optional[float] = None) -> Dict[str, Tensor]:
        N, _, H, W = target.size()
        out = {}
        num_pixels = N * H * W
        if p is None:
            p = self.lmbda[0]
        bits = None
        for likelihoods in output["likelihoods"].values():
            b = rate_bits(likelihoods).sum()
            bits = b if bits is None else bits + b
        out["bpp_loss"] = (bits / num_pixels).to(target.dtype)
        if self.type == "mse":
            out["mse_loss"] = self.mse(output["x_hat"], target)
            out["loss"] = p * 255 ** 2 * out["mse_loss"] + out["bpp_loss"]
        else:
            raise NotImplementedError("RateDistortionLoss(type='ms_ssim') needs pytorch_msssim (utils/helper.compute_msssim); "
                                      "only the rate term and the MSE distortion are part of this path")
        return out
