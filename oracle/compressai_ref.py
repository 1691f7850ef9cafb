"""TEST INFRASTRUCTURE ONLY — CPU restatement of the CompressAI-style entropy path.

``TCM`` (reference ``src/models/reference/tcm.py:416-417``) instantiates
``compressai.entropy_models.{EntropyBottleneck, GaussianConditional}``.  CompressAI is a
third-party dependency that is NOT vendored under ``/root/reference`` and is unpinned
(``Dockerfile:5`` — ``pip install compressai`` on a torch-1.10 image; API usage implies
>= 1.2.x).  This file restates its published algorithm in plain, unfused torch CPU ops,
one reference op per line, anchored on the reference's own copies of the same code:

* base class / buffers ........ ``src/entropy_models/adaptive_gaussian_conditional.py:17-61``
* ``_standardized_cumulative`` . ``adaptive_gaussian_conditional.py:375-379``, ``tcm.py:584-588``
* GC ``_likelihood`` twin ...... ``src/models/reference/tcm.py:570-582``
* ``build_indexes`` ............ ``adaptive_gaussian_conditional.py:606-617``
* GC ``update`` (upstream) ..... ``adaptive_gaussian_conditional.py:457-482`` (kept as comment there)
* ``_pmf_to_cdf`` .............. ``adaptive_gaussian_conditional.py:197-205``
* EB parameters / init ......... ``src/entropy_models/adaptive_entropy_bottleneck.py:341-362,384-385``
* ``_logits_cumulative`` ....... ``adaptive_entropy_bottleneck.py:525-543``
* sign-trick likelihood ........ ``adaptive_entropy_bottleneck.py:658-666``
* permute wrapper .............. ``adaptive_entropy_bottleneck.py:679-708``
* ``ste_round`` / scale table .. ``tcm.py:26-37``
* rate reduction ............... ``src/training/loss.py:24-27``, ``src/eval.py:27-31``

All functions work in the dtype of their inputs (fp32 = the reference; fp64 = "truth").
Parity pinned by ``oracle/gen_golden.py`` -> ``tests/golden`` (see oracle/__init__.py).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.stats
import torch
import torch.nn.functional as F
from torch import Tensor

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64  # tcm.py:26-28


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS) -> Tensor:
    """tcm.py:33-34."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def ste_round(x: Tensor) -> Tensor:
    """tcm.py:36-37 (value == round(x); gradient identity)."""
    return torch.round(x) - x.detach() + x


def lower_bound(x: Tensor, bound: float) -> Tensor:
    """compressai.ops.LowerBound forward: ``torch.max(x, bound)`` with a 1-element
    *tensor* bound of x's dtype (so NaN propagates).  Backward rule:
    pass the gradient where ``x >= bound`` or ``grad < 0``."""
    return torch.max(x, torch.tensor([bound], dtype=x.dtype, device=x.device))


def lower_bound_backward(x: Tensor, bound: float, grad: Tensor) -> Tensor:
    b = torch.tensor([bound], dtype=x.dtype, device=x.device)
    return ((x >= b) | (grad < 0)).to(grad.dtype) * grad


# ----------------------------------------------------------------------------- quantize
def quantize(inputs: Tensor, mode: str, means: Optional[Tensor] = None,
             noise: Optional[Tensor] = None) -> Tensor:
    """CompressAI ``EntropyModel.quantize``.  ``noise`` replaces the
    ``uniform_(-0.5, 0.5)`` draw so that runs are reproducible."""
    if mode not in ("noise", "dequantize", "symbols"):
        raise ValueError(f'Invalid quantization mode: "{mode}"')
    if mode == "noise":
        if noise is None:
            noise = torch.empty_like(inputs).uniform_(-0.5, 0.5)
        return inputs + noise
    outputs = inputs.clone()
    if means is not None:
        outputs -= means
    outputs = torch.round(outputs)
    if mode == "dequantize":
        if means is not None:
            outputs += means
        return outputs
    return outputs.int()


def dequantize(inputs: Tensor, means: Optional[Tensor] = None, dtype=torch.float) -> Tensor:
    if means is not None:
        outputs = inputs.type_as(means)
        outputs += means
    else:
        outputs = inputs.type(dtype)
    return outputs


# ------------------------------------------------------------------ Gaussian conditional
def standardized_cumulative(inputs: Tensor) -> Tensor:
    """adaptive_gaussian_conditional.py:375-379 / tcm.py:584-588."""
    half = float(0.5)
    const = float(-(2 ** -0.5))
    return half * torch.erfc(const * inputs)


def gc_likelihood(inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                  scale_bound: float = 0.11) -> Tensor:
    """tcm.py:570-582 (== CompressAI GaussianConditional._likelihood)."""
    half = float(0.5)
    if means is not None:
        values = inputs - means
    else:
        values = inputs
    scales = lower_bound(scales, scale_bound)
    values = torch.abs(values)
    upper = standardized_cumulative((half - values) / scales)
    lower = standardized_cumulative((-half - values) / scales)
    return upper - lower


def gc_forward(inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
               training: bool = False, noise: Optional[Tensor] = None,
               scale_bound: float = 0.11, likelihood_bound: float = 1e-9
               ) -> Tuple[Tensor, Tensor]:
    """CompressAI GaussianConditional.forward (call site tcm.py:455)."""
    outputs = quantize(inputs, "noise" if training else "dequantize", means, noise=noise)
    likelihood = gc_likelihood(outputs, scales, means, scale_bound)
    if likelihood_bound > 0:
        likelihood = lower_bound(likelihood, likelihood_bound)
    return outputs, likelihood


def build_indexes(scales: Tensor, scale_table: Tensor, scale_bound: float = 0.11) -> Tensor:
    """adaptive_gaussian_conditional.py:606-617 — the 63-pass loop, verbatim order."""
    scales = lower_bound(scales, scale_bound)
    indexes = scales.new_full(scales.size(), len(scale_table) - 1).int()
    for s in scale_table[:-1]:
        indexes -= (scales <= s).int()
    return indexes


def pmf_to_quantized_cdf(pmf: Sequence[float], precision: int = 16) -> List[int]:
    """compressai._CXX.pmf_to_quantized_cdf (C++), restated with Python ints."""
    for p in pmf:
        if p < 0 or not math.isfinite(p):
            raise ValueError(f"Invalid `pmf`, non-finite or negative element found: {p}")
    cdf = [0] * (len(pmf) + 1)
    for i, p in enumerate(pmf):
        # std::round on a float32 product (pmf is std::vector<float>): half away from zero
        v = float(np.float32(p)) * (1 << precision)
        cdf[i + 1] = int(math.floor(v + 0.5))
    total = sum(cdf)
    if total == 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    for i in range(len(cdf)):
        cdf[i] = ((1 << precision) * cdf[i]) // total
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]
    cdf[-1] = 1 << precision
    for i in range(len(cdf) - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq = 0xFFFFFFFF
            best_steal = -1
            for j in range(len(cdf) - 1):
                freq = cdf[j + 1] - cdf[j]
                if freq > 1 and freq < best_freq:
                    best_freq = freq
                    best_steal = j
            assert best_steal != -1
            if best_steal < i:
                for j in range(best_steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                assert best_steal > i
                for j in range(i + 1, best_steal + 1):
                    cdf[j] += 1
    assert cdf[0] == 0 and cdf[-1] == (1 << precision)
    for i in range(len(cdf) - 1):
        assert cdf[i + 1] > cdf[i], "Invalid cdf: not strictly monotonic"
    return cdf


def _pmf_to_cdf(pmf: Tensor, tail_mass: Tensor, pmf_length: Tensor, max_length: int,
                precision: int = 16) -> Tensor:
    """adaptive_gaussian_conditional.py:197-205."""
    cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
    for i, p in enumerate(pmf):
        prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
        _cdf = torch.IntTensor(pmf_to_quantized_cdf(prob.tolist(), precision))
        cdf[i, : _cdf.size(0)] = _cdf
    return cdf


def gc_update(scale_table: Tensor, tail_mass: float = 1e-9, precision: int = 16):
    """CompressAI GaussianConditional.update — the block kept as a comment at
    adaptive_gaussian_conditional.py:457-482.  Returns (_quantized_cdf, _offset, _cdf_length)."""
    multiplier = -scipy.stats.norm.ppf(tail_mass / 2)
    pmf_center = torch.ceil(scale_table * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = torch.max(pmf_length).item()
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None])
    samples_scale = scale_table.unsqueeze(1)
    samples = samples.float()
    samples_scale = samples_scale.float()
    upper = standardized_cumulative((0.5 - samples) / samples_scale)
    lower = standardized_cumulative((-0.5 - samples) / samples_scale)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    quantized_cdf = _pmf_to_cdf(pmf, tail, pmf_length, max_length, precision)
    return quantized_cdf, -pmf_center, pmf_length + 2


# ------------------------------------------------------------------ entropy bottleneck
class EntropyBottleneckRef:
    """CompressAI EntropyBottleneck(channels, tail_mass=1e-9, init_scale=10,
    filters=(3,3,3,3)) restated on plain tensors."""

    def __init__(self, channels: int, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters: Tuple[int, ...] = (3, 3, 3, 3), likelihood_bound: float = 1e-9,
                 generator: Optional[torch.Generator] = None, dtype=torch.float32):
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        self.likelihood_bound = float(likelihood_bound)
        # adaptive_entropy_bottleneck.py:341-362
        filt = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        self.matrices, self.biases, self.factors = [], [], []
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filt[i + 1]))
            matrix = torch.empty(channels, filt[i + 1], filt[i], dtype=dtype).fill_(init)
            self.matrices.append(matrix)
            bias = torch.empty(channels, filt[i + 1], 1, dtype=dtype)
            bias.uniform_(-0.5, 0.5, generator=generator)
            self.biases.append(bias)
            if i < len(self.filters):
                self.factors.append(torch.zeros(channels, filt[i + 1], 1, dtype=dtype))
        self.quantiles = torch.tensor([-self.init_scale, 0.0, self.init_scale], dtype=dtype
                                      ).repeat(channels, 1, 1)  # [C,1,3]
        target = np.log(2 / self.tail_mass - 1)  # :384-385
        self.target = torch.tensor([-target, 0, target], dtype=dtype)

    def to(self, dtype):
        out = EntropyBottleneckRef.__new__(EntropyBottleneckRef)
        out.__dict__ = dict(self.__dict__)
        out.matrices = [m.to(dtype) for m in self.matrices]
        out.biases = [m.to(dtype) for m in self.biases]
        out.factors = [m.to(dtype) for m in self.factors]
        out.quantiles = self.quantiles.to(dtype)
        out.target = self.target.to(dtype)
        return out

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, inputs: Tensor) -> Tensor:
        """adaptive_entropy_bottleneck.py:525-543."""
        logits = inputs
        for i in range(len(self.filters) + 1):
            logits = torch.matmul(F.softplus(self.matrices[i]), logits)
            logits = logits + self.biases[i]
            if i < len(self.filters):
                logits = logits + torch.tanh(self.factors[i]) * torch.tanh(logits)
        return logits

    def _likelihood(self, inputs: Tensor) -> Tensor:
        """Upstream ±0.5 bins + the sign trick of adaptive_entropy_bottleneck.py:658-666."""
        half = float(0.5)
        v0 = inputs - half
        v1 = inputs + half
        lower = self._logits_cumulative(v0)
        upper = self._logits_cumulative(v1)
        sign = -torch.sign(lower + upper)
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def forward(self, x: Tensor, training: bool = False, noise: Optional[Tensor] = None
                ) -> Tuple[Tensor, Tensor]:
        """Permute wrapper of adaptive_entropy_bottleneck.py:679-708 with CompressAI's
        quantize-about-the-medians."""
        perm = list(range(x.dim()))
        perm[0], perm[1] = perm[1], perm[0]
        inv_perm = list(np.argsort(perm))
        xp = x.permute(*perm).contiguous()
        shape = xp.size()
        values = xp.reshape(xp.size(0), 1, -1)
        nz = None
        if noise is not None:
            nz = noise.permute(*perm).contiguous().reshape(xp.size(0), 1, -1)
        outputs = quantize(values, "noise" if training else "dequantize", self._get_medians(),
                           noise=nz)
        likelihood = self._likelihood(outputs)
        if self.likelihood_bound > 0:
            likelihood = lower_bound(likelihood, self.likelihood_bound)
        outputs = outputs.reshape(shape).permute(*inv_perm).contiguous()
        likelihood = likelihood.reshape(shape).permute(*inv_perm).contiguous()
        return outputs, likelihood

    def symbols(self, x: Tensor) -> Tensor:
        """compress() front half: int32(round(x - median)) with the per-channel index."""
        med = self._get_medians().reshape(1, -1, *([1] * (x.dim() - 2)))
        return quantize(x, "symbols", med.expand_as(x))

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles)
        return torch.abs(logits - self.target).sum()

    def update(self, precision: int = 16):
        """CompressAI EntropyBottleneck.update."""
        medians = self.quantiles[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - self.quantiles[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - medians).int(), min=0)
        offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = pmf_length.max().item()
        samples = torch.arange(max_length)
        samples = samples[None, :] + pmf_start[:, None, None]
        half = float(0.5)
        lower = self._logits_cumulative(samples - half)
        upper = self._logits_cumulative(samples + half)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        pmf = pmf[:, 0, :]
        tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        quantized_cdf = _pmf_to_cdf(pmf, tail, pmf_length, max_length, precision)
        return quantized_cdf, offset, pmf_length + 2


# ----------------------------------------------------------------------------- rate
def bpp(likelihoods: Sequence[Tensor], num_pixels: int) -> Tensor:
    """loss.py:24-27 / eval.py:27-31: sum_t log(L_t).sum() / (-ln2 * num_pixels)."""
    return sum((torch.log(l).sum() / (-math.log(2) * num_pixels)) for l in likelihoods)


def per_image_bits(likelihood: Tensor) -> Tensor:
    """Same reduction kept per image (float64 accumulate): bits_b = -sum log2 L[b]."""
    l = likelihood.reshape(likelihood.shape[0], -1)
    return (torch.log(l).double().sum(dim=1) / (-math.log(2)))


# ----------------------------------------------------- full per-slice reference "step"
def tcm_entropy_step(y: Tensor, mu: Tensor, sigma: Tensor, z: Tensor, eb: EntropyBottleneckRef,
                     scale_table: Tensor, num_slices: int = 5, training: bool = False,
                     with_indexes: bool = True, num_pixels: Optional[int] = None,
                     noise_y: Optional[Tensor] = None, noise_z: Optional[Tensor] = None):
    """One pass of the hot path exactly as TCM.forward / TCM.compress drive it
    (tcm.py:429-466 and :527-552), minus the dense networks: the per-slice Gaussian
    conditional + ste_round, optional build_indexes + symbols, the bottleneck on z, bpp."""
    out = {}
    _, z_lik = eb.forward(z, training=training, noise=noise_z)
    med = eb._get_medians().reshape(1, -1, 1, 1)
    out["z_hat"] = ste_round(z - med) + med          # tcm.py:431-433
    y_hat, y_lik, idx, sym = [], [], [], []
    ys, mus, sgs = y.chunk(num_slices, 1), mu.chunk(num_slices, 1), sigma.chunk(num_slices, 1)
    nzs = noise_y.chunk(num_slices, 1) if noise_y is not None else [None] * num_slices
    for y_s, m_s, s_s, n_s in zip(ys, mus, sgs, nzs):
        _, lik = gc_forward(y_s, s_s, m_s, training=training, noise=n_s)   # tcm.py:455
        y_lik.append(lik)
        y_hat.append(ste_round(y_s - m_s) + m_s)                           # tcm.py:457
        if with_indexes:
            idx.append(build_indexes(s_s, scale_table))                    # tcm.py:544
            sym.append(quantize(y_s, "symbols", m_s))                      # tcm.py:548
    out["y_hat"] = torch.cat(y_hat, 1)
    out["y_lik"] = torch.cat(y_lik, 1)
    out["z_lik"] = z_lik
    if with_indexes:
        out["indexes"] = torch.cat(idx, 1)
        out["symbols"] = torch.cat(sym, 1)
    if num_pixels is None:
        num_pixels = y.shape[0] * y.shape[2] * 16 * y.shape[3] * 16
    out["bpp"] = bpp([out["y_lik"], z_lik], num_pixels)
    return out
