"""Drop-in entropy-model modules for ResLIC_TCM, backed by fused sm_100a kernels.

Host-side mirror of the operator API the reference's models call
(SURVEY.md §8b): ``compressai.entropy_models.{EntropyModel, GaussianConditional,
EntropyBottleneck}`` and ``compressai.ops.LowerBound`` — same constructor signatures,
method names, argument meaning, error behaviour, parameter and buffer names/shapes, so
``TCM.forward/compress/decompress`` (src/models/reference/tcm.py:425-635) and reference
checkpoints work unchanged:

    self.entropy_bottleneck = EntropyBottleneck(192)        # tcm.py:416
    self.gaussian_conditional = GaussianConditional(None)   # tcm.py:417

Per-element work (quantize, likelihood, bounds, index build, rate sum) runs in ONE fused
CUDA launch per call through the C ABI; there is no PyTorch or CPU fallback for it.
Setup code that runs once per ``update()`` (CDF tables) stays in torch + a host C++ helper.

Extra, non-reference API (used by the fused TCM slice loop and the benchmark):
``GaussianConditional.forward_fused`` / ``EntropyBottleneck.forward_fused`` return every
by-product of the pass (ste_round output, symbols, indexes, per-image bits) from the same
launch.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.stats
import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from ._cabi import ReslicError

__all__ = ["LowerBound", "EntropyModel", "GaussianConditional", "EntropyBottleneck"]


class _LowerBoundFunction(torch.autograd.Function):
    """max(x, bound) whose gradient passes where x >= bound or the step moves x up
    (compressai.ops.LowerBound)."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        keep = (x >= bound) | (grad_output < 0)
        return keep * grad_output, None


class LowerBound(nn.Module):
    """``compressai.ops.LowerBound`` (used standalone by reference code, e.g.
    adaptive_gaussian_conditional.py:35,343).  Inside the fused kernels the bounds are
    applied in-register; this module exists for API compatibility."""

    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFunction.apply(x, self.bound)


def _no_grad_path(*tensors):
    """Guard of the few entry points that have no backward kernel (the fused multi-output ``forward_fused`` forms
    and ``EntropyBottleneckStanh.forward``): refuse rather than silently cut the graph.  The module ``forward``s, the
    ``quantize`` entry points and the STanH activation record autograd nodes (backward kernels: gc_bwd.cu, eb_bwd.cu,
    stanh_fused.cu)."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise ReslicError(
            "this reslic_tcm_b200 entry point has no backward kernel: call it under torch.no_grad() "
            "(differentiable forms: the modules' forward() / quantize())"
        )


class _QuantizeFn(torch.autograd.Function):
    """``EntropyModel.quantize`` with the gradient autograd gives the reference:
    "noise": out = inputs + U(-1/2, 1/2) -> identity to inputs (means ignored); "dequantize":
    out = round(inputs - means) + means -> zero to inputs (torch.round), identity to means."""

    @staticmethod
    def forward(ctx, inputs, means, mode, seed, offset):
        ctx.mode, ctx.has_means = mode, means is not None
        if mode == "noise":
            return ops.gc_forward(inputs, None, None, training=True, want=("yhat",), seed=seed, offset=offset).yhat
        return ops.gc_forward(inputs, None, means, want=("ste",)).ste

    @staticmethod
    def backward(ctx, g):
        if ctx.mode == "noise":
            return g, None, None, None, None
        g_in = torch.zeros_like(g) if ctx.needs_input_grad[0] else None
        return g_in, (g if ctx.has_means and ctx.needs_input_grad[1] else None), None, None, None


class _GaussianConditionalFn(torch.autograd.Function):
    """Autograd node around the fused forward / backward kernels (one launch each)."""

    @staticmethod
    def forward(ctx, inputs, scales, means, noise, training, scale_bound, lik_bound, seed, offset):
        r = ops.gc_forward(inputs, scales, means, training=training, noise=noise, scale_bound=scale_bound,
                           likelihood_bound=lik_bound, want=("yhat", "lik", "ste"), seed=seed, offset=offset)
        ctx.save_for_backward(inputs, scales, means, noise)
        ctx.cfg = (bool(training), float(scale_bound), float(lik_bound), int(seed), int(offset))
        ctx.set_materialize_grads(False)
        return r.yhat, r.lik, r.ste

    @staticmethod
    def backward(ctx, g_yhat, g_lik, g_ste):
        inputs, scales, means, noise = ctx.saved_tensors
        training, scale_bound, lik_bound, seed, offset = ctx.cfg
        need = (ctx.needs_input_grad[0], means is not None and ctx.needs_input_grad[2], ctx.needs_input_grad[1])
        c = lambda t: None if t is None else t.contiguous()
        g_y, g_mu, g_sigma = ops.gc_backward(inputs, scales, means, training=training, noise=noise,
                                             scale_bound=scale_bound, likelihood_bound=lik_bound, g_yhat=c(g_yhat),
                                             g_ste=c(g_ste), g_lik=c(g_lik), need=need, seed=seed, offset=offset)
        return g_y, g_sigma, g_mu, None, None, None, None, None, None


class _EntropyBottleneckFn(torch.autograd.Function):
    """Autograd node around the fused bottleneck forward / backward kernels."""

    @staticmethod
    def forward(ctx, x, noise, training, lik_bound, seed, offset, quantiles, *params):
        m, b, f = params[:5], params[5:10], params[10:14]
        med = quantiles.detach()[:, 0, 1].contiguous()
        r = ops.eb_forward(x, m, b, f, med, training=training, noise=noise, likelihood_bound=lik_bound,
                           want=("zhat", "lik"), seed=seed, offset=offset)
        ctx.save_for_backward(x, noise, quantiles, *params)
        ctx.cfg = (bool(training), float(lik_bound), int(seed), int(offset))
        ctx.set_materialize_grads(False)
        return r.zhat, r.lik

    @staticmethod
    def backward(ctx, g_zhat, g_lik):
        x, noise, quantiles, *params = ctx.saved_tensors
        training, lik_bound, seed, offset = ctx.cfg
        m, b, f = params[:5], params[5:10], params[10:14]
        c = lambda t: None if t is None else t.contiguous()
        g_z, gm, gb, gf, g_med = ops.eb_backward(
            x, m, b, f, quantiles.detach()[:, 0, 1], training=training, noise=noise, likelihood_bound=lik_bound,
            g_zhat=c(g_zhat), g_lik=c(g_lik), need_z=ctx.needs_input_grad[0],
            need_params=any(ctx.needs_input_grad[7:]), seed=seed, offset=offset)
        g_q = None
        if ctx.needs_input_grad[6]:
            g_q = torch.zeros_like(quantiles)
            g_q[:, 0, 1] = g_med
        gp = (list(gm) + list(gb) + list(gf)) if gm else [None] * 14
        return (g_z, None, None, None, None, None, g_q, *gp)


class EntropyModel(nn.Module):
    """Base class: CDF buffers + quantize/dequantize (reference copy of the upstream
    class: src/entropy_models/adaptive_gaussian_conditional.py:17-61)."""

    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder_name = entropy_coder or "ans"
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        self._likelihood_bound = float(likelihood_bound)
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        # filled by update()
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    def forward(self, *args: Any) -> Any:
        raise NotImplementedError()

    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        """modes "noise" | "dequantize" | "symbols" (SURVEY.md App. A.1)."""
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "symbols":       # integers: nothing to differentiate
            return ops.gc_forward(inputs.detach(), None, None if means is None else means.detach(), want=("sym",)).sym
        seed, offset = _philox_state(inputs.numel()) if mode == "noise" else (0, 0)
        if means is not None and means.shape != inputs.shape:
            means = means.expand_as(inputs)
        # the reference calls this with grad enabled (GainBalle2018: quantize(y, "noise") inside the training forward)
        return _QuantizeFn.apply(inputs, None if mode == "noise" else means, mode, seed, offset)

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None, dtype: torch.dtype = torch.float) -> Tensor:
        if means is not None:
            if inputs.dtype == torch.int32 and means.dtype == torch.float32 and inputs.is_cuda:
                return ops.dequantize(inputs, means)
            outputs = inputs.type_as(means)
            outputs += means
            return outputs
        return inputs.type(dtype)

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        """adaptive_gaussian_conditional.py:197-205, rows quantised by the host C++ helper."""
        from . import cdf_tables

        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32, device=pmf.device)
        pmf_c, tail_c, len_c = pmf.detach().cpu(), tail_mass.detach().cpu(), pmf_length.detach().cpu()
        for i in range(len(len_c)):
            prob = torch.cat((pmf_c[i, : int(len_c[i])], tail_c[i].reshape(-1)), dim=0)
            row = cdf_tables.pmf_to_quantized_cdf(prob, self.entropy_coder_precision)
            cdf[i, : row.numel()] = row.to(cdf.device)
        return cdf

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def compress(self, inputs, indexes, means=None):
        """Symbols + indexes come from the fused kernel (one D2H copy each, no .tolist()); the strings
        from the host rANS coder, one per image, images coded on a thread pool."""
        symbols = self.quantize(inputs, "symbols", means)
        if len(inputs.size()) < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        if symbols.is_cuda and indexes.is_cuda and indexes.dtype == torch.int32 and self._quantized_cdf.size(0) <= 1024:
            # table lookups on the device, one packed slot per symbol over PCIe; same strings as the host lookup
            out = ops.rans_slots(symbols, indexes, self._quantized_cdf, self._cdf_length, self._offset)
            strings = _rans().encode_slots_batch(*out)
            if strings is not None:
                return strings
        return _rans().encode_with_indexes_batch(symbols, indexes, self._quantized_cdf, self._cdf_length,
                                                 self._offset)

    def decompress(self, strings, indexes, dtype: torch.dtype = torch.float, means: Optional[Tensor] = None):
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, len(indexes.size())):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        symbols = _rans().decode_with_indexes_batch(strings, indexes, self._quantized_cdf, self._cdf_length,
                                                    self._offset)
        return self.dequantize(symbols, means, dtype)


def _rans():
    """The host rANS coder (csrc/rans.cpp behind reslic_tcm_b200.rans — compressai.ans' API); its
    inputs, the symbols and CDF indexes (or packed slots), are produced on the GPU by this package."""
    from . import rans

    return rans


_philox_counter = [0]


def _philox_state(numel: int) -> Tuple[int, int]:
    """(seed, offset) for the in-kernel Philox stream: seed follows torch's CUDA generator
    (so ``torch.manual_seed`` makes runs reproducible), offset advances per call."""
    seed = torch.cuda.initial_seed() if torch.cuda.is_available() else torch.initial_seed()
    off = _philox_counter[0]
    _philox_counter[0] += 1
    return seed, off


class GaussianConditional(EntropyModel):
    """``compressai.entropy_models.GaussianConditional`` (STanH copy of
    the constructor: adaptive_gaussian_conditional.py:314-353)."""

    def __init__(self, scale_table: Optional[Union[List, Tuple]], *args: Any, scale_bound: float = 0.11,
                 tail_mass: float = 1e-9, **kwargs: Any):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = self.scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self._scale_bound = float(scale_bound)
        self.register_buffer("scale_table",
                             self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound",
                             torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_cumulative(self, inputs: Tensor) -> Tensor:
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force: bool = False) -> bool:
        # offsets are only computed when the conditional model is updated
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        """CDF tables for the rANS coder (setup, once per model; upstream algorithm kept as
        a comment at adaptive_gaussian_conditional.py:457-482)."""
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(self.scale_table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = torch.max(pmf_length).item()
        device = pmf_center.device
        samples = torch.abs(torch.arange(max_length, device=device).int() - pmf_center[:, None])
        samples_scale = self.scale_table.unsqueeze(1)
        samples = samples.float()
        samples_scale = samples_scale.float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._offset = -pmf_center
        self._cdf_length = pmf_length + 2

    # ---- per-element path: fused kernel -------------------------------------------------
    def forward_fused(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                      training: Optional[bool] = None, want: Sequence[str] = ("yhat", "lik"),
                      noise: Optional[Tensor] = None, out: Optional[dict] = None) -> ops.GcOutputs:
        if training is None:
            training = self.training
        _no_grad_path(inputs, scales, means)
        seed, offset = (0, 0)
        if training and noise is None:
            seed, offset = _philox_state(inputs.numel())
        return ops.gc_forward(
            inputs, scales, means, training=training, noise=noise,
            scale_table=self.scale_table if "idx" in want else None,
            scale_bound=self._scale_bound,
            likelihood_bound=self._likelihood_bound if self.use_likelihood_bound else 0.0,
            want=want, out=out, seed=seed, offset=offset)

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (inputs, scales, means)):
            y_hat, lik, _ = self.forward_with_ste(inputs, scales, means, training)
            return y_hat, lik
        r = self.forward_fused(inputs, scales, means, training, want=("yhat", "lik"))
        return r.yhat, r.lik

    def forward_with_ste(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                         training: Optional[bool] = None, noise: Optional[Tensor] = None
                         ) -> Tuple[Tensor, Tensor, Tensor]:
        """Differentiable (y_hat, likelihood, ste_round(y - mu) + mu) from one launch: the pair of
        calls at tcm.py:455 and :457, with the fused backward kernel behind autograd."""
        if training is None:
            training = self.training
        seed, offset = (0, 0)
        if training and noise is None:
            seed, offset = _philox_state(inputs.numel())
        return _GaussianConditionalFn.apply(
            inputs, scales, means, noise, bool(training), self._scale_bound,
            self._likelihood_bound if self.use_likelihood_bound else 0.0, seed, offset)

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        """Unbounded likelihood of already-quantised ``inputs`` (tcm.py:570-582).  Served by
        the same kernel: feeding ŷ as y in NOISE mode with zero noise reproduces
        v = |ŷ - mu| exactly."""
        _no_grad_path(inputs, scales, means)
        zero = torch.zeros_like(inputs)
        return ops.gc_forward(inputs, scales, means, training=True, noise=zero,
                              scale_bound=self._scale_bound, likelihood_bound=0.0, want=("lik",)).lik

    def build_indexes(self, scales: Tensor) -> Tensor:
        """adaptive_gaussian_conditional.py:606-617 — one launch instead of ~192."""
        if self.scale_table.numel() == 0:
            raise ValueError("Uninitialized scale_table. Run update_scale_table() first")
        return ops.build_indexes(scales, self.scale_table, self._scale_bound)


class EntropyBottleneck(EntropyModel):
    """``compressai.entropy_models.EntropyBottleneck`` (parameters as in
    src/entropy_models/adaptive_entropy_bottleneck.py:341-362)."""

    _offset: Tensor

    def __init__(self, channels: int, *args: Any, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters: Tuple[int, ...] = (3, 3, 3, 3), **kwargs: Any):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self) -> Tensor:
        medians = self.quantiles[:, :, 1:2]
        return medians

    def _medians_flat(self) -> Tensor:
        """Contiguous [C] copy of the medians for the kernel, refreshed only when `quantiles`
        changes (keeps a strided-copy kernel out of every forward / CUDA graph)."""
        q = self.quantiles
        key = (id(q), q._version, q.data_ptr(), q.device)
        if getattr(self, "_med_key", None) != key:
            self._med_flat = q.detach()[:, 0, 1].contiguous()
            self._med_key = key
        return self._med_flat

    def _eval_lut(self) -> Tensor:
        """[C, 130] table of the eval-mode likelihoods (ops.eb_build_lut), rebuilt only when a
        parameter, the medians or the bound changes (tensor version counters): evaluation and
        compress loops then launch the bottleneck without its 5-layer table construction."""
        m, b, f = self._params()
        bound = self._likelihood_bound if self.use_likelihood_bound else 0.0
        key = tuple((id(t), t._version, t.data_ptr()) for t in (*m, *b, *f, self.quantiles)) + (bound,)
        if getattr(self, "_lut_key", None) != key:
            self._lut = ops.eb_build_lut(m, b, f, self._medians_flat(), likelihood_bound=bound)
            self._lut_key = key
        return self._lut

    def _params(self):
        n = len(self.filters) + 1
        return ([getattr(self, f"_matrix{i:d}") for i in range(n)],
                [getattr(self, f"_bias{i:d}") for i in range(n)],
                [getattr(self, f"_factor{i:d}") for i in range(n - 1)])

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        """Setup-time evaluation on [C,1,L] sample grids (update(), loss()); the per-element
        evaluation on z lives in the fused kernel (adaptive_entropy_bottleneck.py:525-543)."""
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            if stop_gradient:
                matrix = matrix.detach()
            logits = torch.matmul(torch.nn.functional.softplus(matrix), logits)
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                bias = bias.detach()
            logits = logits + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        loss = torch.abs(logits - self.target).sum()
        return loss

    def update(self, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        medians = self.quantiles[:, 0, 1]
        minima = medians - self.quantiles[:, 0, 0]
        minima = torch.ceil(minima).int()
        minima = torch.clamp(minima, min=0)
        maxima = self.quantiles[:, 0, 2] - medians
        maxima = torch.ceil(maxima).int()
        maxima = torch.clamp(maxima, min=0)
        self._offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = pmf_length.max().item()
        device = pmf_start.device
        samples = torch.arange(max_length, device=device)
        samples = samples[None, :] + pmf_start[:, None, None]
        half = float(0.5)
        lower = self._logits_cumulative(samples - half, stop_gradient=True)
        upper = self._logits_cumulative(samples + half, stop_gradient=True)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        pmf = pmf[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._quantized_cdf = quantized_cdf
        self._cdf_length = pmf_length + 2
        return True

    # ---- per-element path: fused kernel -------------------------------------------------
    def forward_fused(self, x: Tensor, training: Optional[bool] = None,
                      want: Sequence[str] = ("zhat", "lik"), noise: Optional[Tensor] = None) -> ops.EbOutputs:
        if training is None:
            training = self.training
        if self.filters != (3, 3, 3, 3):
            raise ReslicError("the CUDA bottleneck supports filters=(3,3,3,3) only")
        m, b, f = self._params()
        _no_grad_path(x, *m, *b, *f, self.quantiles)
        seed, offset = (0, 0)
        if training and noise is None:
            seed, offset = _philox_state(x.numel())
        return ops.eb_forward(
            x, m, b, f, self._medians_flat(), training=training, noise=noise,
            likelihood_bound=self._likelihood_bound if self.use_likelihood_bound else 0.0,
            want=want, seed=seed, offset=offset, lut=None if training else self._eval_lut())

    def forward(self, x: Tensor, training: Optional[bool] = None, noise: Optional[Tensor] = None
                ) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training
        m, b, f = self._params()
        if torch.is_grad_enabled() and any(t.requires_grad for t in (x, self.quantiles, *m, *b, *f)):
            if self.filters != (3, 3, 3, 3):
                raise ReslicError("the CUDA bottleneck supports filters=(3,3,3,3) only")
            seed, offset = (0, 0)
            if training and noise is None:
                seed, offset = _philox_state(x.numel())
            return _EntropyBottleneckFn.apply(
                x, noise, bool(training), self._likelihood_bound if self.use_likelihood_bound else 0.0, seed, offset,
                self.quantiles, *m, *b, *f)
        r = self.forward_fused(x, training, want=("zhat", "lik"), noise=noise)
        return r.zhat, r.lik

    @staticmethod
    def _build_indexes(size):
        dims = len(size)
        N = size[0]
        C = size[1]
        view_dims = np.ones((dims,), dtype=np.int64)
        view_dims[1] = -1
        indexes = torch.arange(C).view(*view_dims)
        indexes = indexes.int()
        return indexes.repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def compress(self, x):
        indexes = self._build_indexes(x.size()).to(x.device)
        medians = self._get_medians().detach()
        spatial_dims = len(x.size()) - 2
        medians = self._extend_ndims(medians, spatial_dims)
        medians = medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))
        return super().compress(x, indexes, medians)

    def decompress(self, strings, size):
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        indexes = self._build_indexes(output_size).to(self._quantized_cdf.device)
        medians = self._extend_ndims(self._get_medians().detach(), len(size))
        medians = medians.expand(len(strings), *([-1] * (len(size) + 1)))
        return super().decompress(strings, indexes, medians.dtype, medians)
