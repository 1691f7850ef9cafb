"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's STanH path, op for op.

Follows src/quantization/activation.py (NonSymStanH :7-150, SymStanH :157-304) and
src/entropy_models/adaptive_gaussian_conditional.py (quantize :95-157, define_v0_and_v1 :495-537,
_likelihood :541-580, forward :588-603) and compute_gap (src/models/stanh/tcm_stanh.py:465-478).
Pinned against the reference's own modules by tests/golden/stanh_golden.npz
(oracle/gen_golden.py:gen_stanh)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .compressai_ref import lower_bound, standardized_cumulative


def levels_nonsym(w: Tensor) -> Tensor:
    """activation.py:91-98."""
    n = (torch.sum(w) / 2).item()
    cum_w = torch.zeros(w.numel() + 1)
    cum_w[1:] = torch.cumsum(w, dim=0)
    return torch.sub(cum_w, n)


def levels_sym(w: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """activation.py:214-234 -> (cum_w, sym_w, sym_b is built by the caller)."""
    cum_w = torch.zeros(w.numel() + 1)
    cum_w[1:] = torch.cumsum(w, dim=0)
    return torch.cat((-torch.flip(cum_w[1:], dims=[0]), cum_w), dim=0)


def mid_and_half_gaps(cum_w: Tensor) -> Tuple[Tensor, Tensor]:
    """activation.py:82-88."""
    return torch.add(cum_w[1:], cum_w[:-1]) / 2, torch.sub(cum_w[1:], cum_w[:-1]) / 2


def f(x: Tensor) -> Tensor:
    return 2 * torch.sigmoid(2 * x) - 1


def stanh(x: Tensor, w: Tensor, b_sorted: Tensor, beta: float, symmetric: bool) -> Tensor:
    """The dense [1, K, N] evaluation of activation.py:135-150 / 294-304.  `w` pairs with the
    SORTED thresholds (sym_w / sym_b for the symmetric module)."""
    shape = x.shape
    xx = x.reshape(1, 1, -1)
    b = b_sorted
    if beta == -1:
        if symmetric:
            out = torch.sum((w[:, None] / 2) * torch.sign(xx - b[:, None]), dim=1).unsqueeze(1)
        else:
            out = torch.sum(w[:, None] * torch.relu(torch.sign(xx - b[:, None])) - w[:, None] / 2, dim=1).unsqueeze(1)
    else:
        out = torch.sum((w[:, None] / 2) * f(beta * (xx - b[:, None])), dim=1).unsqueeze(1)
    return out.reshape(shape)


def define_v0_and_v1(inputs: Tensor, average_points: Tensor, distance_points: Tensor) -> Tuple[Tensor, Tensor]:
    """adaptive_gaussian_conditional.py:495-537, verbatim structure (one-hot matrices)."""
    shape = inputs.shape
    x = inputs.reshape(-1).unsqueeze(1)
    left = torch.zeros(average_points.shape[0] + 1) - 1000
    left[1:] = average_points
    right = torch.zeros(average_points.shape[0] + 1) + 1000
    right[:-1] = average_points
    d_left = torch.cat((torch.tensor([0.0]), distance_points), dim=-1).unsqueeze(0)
    d_right = torch.cat((distance_points, torch.tensor([0.0])), dim=-1).unsqueeze(0)
    one_hot = torch.logical_and(x > left.unsqueeze(0), x <= right.unsqueeze(0))
    v0 = torch.sum(d_left * one_hot, dim=1).reshape(shape)
    v1 = torch.sum(d_right * one_hot, dim=1).reshape(shape)
    return v0, v1


def likelihood(inputs: Tensor, scales: Tensor, means: Optional[Tensor], average_points: Tensor,
               distance_points: Tensor, scale_bound: float = 0.11) -> Tensor:
    """adaptive_gaussian_conditional.py:541-580 (works in the dtype of `inputs`)."""
    values = inputs - means if means is not None else inputs
    low, up = define_v0_and_v1(values.float(), average_points, distance_points)
    low, up = low.to(values.dtype), up.to(values.dtype)
    scales = lower_bound(scales, scale_bound)
    upper_pos = standardized_cumulative((low - values) / scales) * (values >= 0)
    upper_neg = standardized_cumulative((values + up) / scales) * (values < 0)
    lower_pos = standardized_cumulative((-up - values) / scales) * (values >= 0)
    lower_neg = standardized_cumulative((values - low) / scales) * (values < 0)
    return (upper_pos + upper_neg) - (lower_pos + lower_neg)


def quantize(inputs: Tensor, mode: str, means: Optional[Tensor], w: Tensor, b_sorted: Tensor, beta: float,
             symmetric: bool, removing_mean: bool) -> Tensor:
    """adaptive_gaussian_conditional.py:95-141 ("training" / "dequantize")."""
    if mode == "training":
        x = inputs - means if (means is not None and removing_mean) else inputs
        out = stanh(x, w, b_sorted, beta, symmetric)
        return out + means if (means is not None and removing_mean) else out
    assert mode == "dequantize"
    out = inputs.clone()
    if means is not None:
        out -= means
    out = stanh(out, w, b_sorted, -1, symmetric)
    if means is not None:
        out += means
    return out


def symbols(inputs: Tensor, means: Optional[Tensor], cum_w: Tensor, w: Tensor, b_sorted: Tensor, symmetric: bool
            ) -> Tensor:
    """"symbols" mode (:144-157): index of the hard level in cum_w (the reference's map_sos_cdf),
    offset so that the symmetric form is centred on 0 (activation.py:252-260)."""
    x = inputs - means if means is not None else inputs
    levels = stanh(x, w, b_sorted, -1, symmetric)
    idx = torch.argmin((levels.reshape(-1, 1) - cum_w.reshape(1, -1)).abs(), dim=1).reshape(inputs.shape)
    off = -(cum_w.numel() // 2) if symmetric else 0
    return (idx + off).int()


def forward(values: Tensor, scales: Tensor, means: Optional[Tensor], training: bool, w: Tensor, b_sorted: Tensor,
            cum_w: Tensor, beta: float, symmetric: bool, removing_mean: bool, likelihood_bound: float = 1e-9
            ) -> Tuple[Tensor, Tensor]:
    """adaptive_gaussian_conditional.py:588-603."""
    avg, dist = mid_and_half_gaps(cum_w)
    y_hat = quantize(values, "training" if training else "dequantize", means, w, b_sorted, beta, symmetric,
                     removing_mean)
    lik = likelihood(y_hat, scales, means, avg, dist)
    if likelihood_bound > 0:
        lik = lower_bound(lik, likelihood_bound)
    return y_hat, lik


def gap(y: Tensor, w: Tensor, b_sorted: Tensor, beta: float, symmetric: bool) -> Tensor:
    """tcm_stanh.py:465-478."""
    import torch.nn.functional as F

    flat = y.reshape(1, 1, -1)
    return torch.abs(F.mse_loss(flat, stanh(flat, w, b_sorted, beta, symmetric)) -
                     F.mse_loss(flat, stanh(flat, w, b_sorted, -1, symmetric)))


def eb_stanh_forward(x: Tensor, eb, w: Tensor, b_sorted: Tensor, cum_w: Tensor, beta: float, symmetric: bool,
                     training: bool, likelihood_bound: float = 1e-9) -> Tuple[Tensor, Tensor]:
    """EntropyBottleneckStanh.forward (src/entropy_models/adaptive_entropy_bottleneck.py:679-708):
    STanH quantization of z WITHOUT medians (EntropyModelSoS.quantize :113-177 with means=None), bins
    [x - low, x + up] from define_v0_and_v1 (:551-603), the same cumulative-logit MLP and sign trick.
    `eb` is a compressai_ref.EntropyBottleneckRef holding the parameters."""
    avg, dist = mid_and_half_gaps(cum_w)
    C = x.shape[1]
    perm = list(range(x.dim()))
    perm[0], perm[1] = perm[1], perm[0]
    xp = x.permute(*perm).contiguous()
    shape = xp.size()
    values = xp.reshape(C, 1, -1)
    flat = values.reshape(1, 1, -1)
    out = stanh(flat, w, b_sorted, beta if training else -1, symmetric).reshape(C, 1, -1)
    low, up = define_v0_and_v1(out.reshape(-1), avg, dist)
    v0 = (out.reshape(-1) - low).reshape(C, 1, -1)
    v1 = (out.reshape(-1) + up).reshape(C, 1, -1)
    lower = eb._logits_cumulative(v0)
    upper = eb._logits_cumulative(v1)
    sign = -torch.sign(lower + upper)
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    if likelihood_bound > 0:
        lik = lower_bound(lik, likelihood_bound)
    inv = [perm.index(i) for i in range(len(perm))]
    return out.reshape(shape).permute(*inv).contiguous(), lik.reshape(shape).permute(*inv).contiguous()
