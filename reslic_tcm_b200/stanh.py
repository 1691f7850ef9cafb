"""STanH ("sum of tanh") quantizers and the STanH Gaussian-conditional entropy model, backed
by the fused sm_100a kernel in ``csrc/stanh_fused.cu``.

Host-side mirror of the reference's in-repo modules (same class names, constructor
arguments, attribute names and method names, so the STanH models call them unchanged):

* ``NonSymStanH`` / ``SymStanH`` ........ src/quantization/activation.py:7-150 / 157-304
* ``GaussianConditionalStanh`` ......... src/entropy_models/adaptive_gaussian_conditional.py:312-725
  (base class ``HypeEntropyModelSoS`` :17-300); call sites src/models/stanh/tcm_stanh.py:331-337,432,
  wacnn_stanh.py:166-171,305, balle18_stanh.py:31-34,126
* ``compute_gap`` ...................... src/models/stanh/tcm_stanh.py:465-478

Reference defects that are NOT replicated: debug prints in hot functions,
the hard-wired ``device=cuda`` defaults, the per-element Python loops of the "symbols" mode
and of ``dequantize``; semantics and API are kept.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, List, Optional, Tuple, Union

import numpy as np
import scipy.stats
import torch
import torch.nn as nn
from torch import Tensor

from . import _cabi, ops
from .entropy_models import EntropyModel, LowerBound, _no_grad_path

__all__ = ["NonSymStanH", "SymStanH", "GaussianConditionalStanh", "EntropyBottleneckStanh", "compute_gap"]


# ----------------------------------------------------------------------------- activation
class _StanHBase(nn.Module):
    symmetric = False

    def _tables(self, beta: float):
        """(struct reslic_stanh_tables, keepalive tensors) for the current state."""
        dev = self.w.device
        key = (id(self.w), id(self.b), self.w._version, self.b._version, self.w.data_ptr(), self.b.data_ptr(), id(self.cum_w), dev)
        cached = getattr(self, "_tables_cache", None)
        if cached is not None and cached[0] == key:
            b_sorted, w, cw, avg, dist = cached[1]
            t = _cabi.StanhTables(b_sorted.data_ptr(), w.data_ptr(), cw.data_ptr(), avg.data_ptr(), dist.data_ptr(),
                                  b_sorted.numel(), 1 if self.symmetric else 0, float(beta))
            return t, cached[1]
        b_sorted = torch.sort(self._all_b().detach())[0].to(dev, torch.float32).contiguous()
        w = self._all_w().detach().to(dev, torch.float32).contiguous()
        cw = self.cum_w.detach().to(dev, torch.float32).contiguous()
        avg = self.average_points.detach().to(dev, torch.float32).contiguous()
        dist = self.distance_points.detach().to(dev, torch.float32).contiguous()
        t = _cabi.StanhTables(b_sorted.data_ptr(), w.data_ptr(), cw.data_ptr(), avg.data_ptr(), dist.data_ptr(),
                              b_sorted.numel(), 1 if self.symmetric else 0, float(beta))
        self._tables_cache = (key, (b_sorted, w, cw, avg, dist))
        return t, (b_sorted, w, cw, avg, dist)

    def _effective(self, w: Tensor, b: Tensor):
        """(w as paired with the sorted thresholds, sorted thresholds, distance_points) as DIFFERENTIABLE
        functions of the raw parameters — the K-sized host mirror of update_state, used only to map the
        kernel's parameter-gradient sums back to w / b."""
        raise NotImplementedError

    def calculate_average_points(self):
        self.average_points = torch.add(self.cum_w[1:], self.cum_w[:-1]) / 2

    def calculate_distance_points(self):
        self.distance_points = torch.sub(self.cum_w[1:], self.cum_w[:-1]) / 2

    def f(self, x):
        return 2 * torch.sigmoid(2 * x) - 1

    def forward(self, x: Tensor, beta=None) -> Tensor:
        """activation.py:135-150 / 294-304: hard levels for beta == -1, else sum of tanh."""
        if beta is None:
            beta = self.beta
        if torch.is_grad_enabled() and any(t.requires_grad for t in (x, self.w, self.b)):
            # the reference evaluates this under grad (update_state / quantize("training"), tcm_stanh.py:399,448)
            return _StanhQuantizeFn.apply(self, x, None, float(beta) != -1.0, False, float(beta), self.w, self.b)
        return stanh_activation(self, x, beta)

    def gap_sums(self, x: Tensor, beta=None, out: Optional[Tensor] = None, workspace: Optional[Tensor] = None) -> Tensor:
        """[sum (x - stanh_beta(x))^2, sum (x - stanh_hard(x))^2] as a float64 GPU tensor.  ``out`` (2 float64) and
        ``workspace`` (>= reslic_stanh_gap_workspace_bytes() zeroed bytes) make the launch allocation-free for callers
        with static buffers (pipeline.TcmStanhEntropyPath)."""
        if beta is None:
            beta = self.beta
        return _stanh_act(self, x, beta, want_soft=False, want_hard=False, want_gap=True, gap_out=out, workspace=workspace)[2]


def _stanh_act(mod: _StanHBase, x: Tensor, beta, want_soft: bool, want_hard: bool, want_gap: bool,
               gap_out: Optional[Tensor] = None, workspace: Optional[Tensor] = None):
    lib = _cabi.load()
    ops._require_cuda("x", x)
    xc = x.contiguous()
    t, keep = mod._tables(beta)
    soft = torch.empty_like(xc) if want_soft else None
    hard = torch.empty_like(xc) if want_hard else None
    gap = ws = None
    if want_gap:
        if gap_out is not None:
            if gap_out.shape != (2,) or gap_out.dtype != torch.float64 or gap_out.device != x.device or not gap_out.is_contiguous():
                raise ValueError("out must be a contiguous float64 tensor of 2 elements on the inputs' device")
            gap = gap_out
        else:
            gap = torch.empty(2, dtype=torch.float64, device=x.device)
        ws = workspace if workspace is not None else _gap_workspace(x.device)
    with torch.cuda.device(x.device):
        code = lib.reslic_stanh_act_f32(xc.data_ptr(), xc.numel(), C.byref(t), _cabi.ptr(soft), _cabi.ptr(hard),
                                        _cabi.ptr(gap), _cabi.ptr(ws), 0 if ws is None else ws.numel(),
                                        _cabi.current_stream_ptr(x.device))
    _cabi.check(code, "reslic_stanh_act_f32")
    del keep
    return soft, hard, gap


_gap_ws = {}


def _gap_workspace(device: torch.device) -> Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    if key not in _gap_ws:
        _gap_ws[key] = torch.zeros(int(_cabi.load().reslic_stanh_gap_workspace_bytes()), dtype=torch.uint8, device=device)
    return _gap_ws[key]


def stanh_activation(mod: _StanHBase, x: Tensor, beta) -> Tensor:
    if beta == -1:
        return _stanh_act(mod, x, -1.0, False, True, False)[1].view_as(x)
    return _stanh_act(mod, x, float(beta), True, False, False)[0].view_as(x)


class NonSymStanH(_StanHBase):
    """activation.py:7-150."""

    symmetric = False

    def __init__(self, beta, num_sigmoids, extrema=5, trainable=True):
        super().__init__()
        self.num_sigmoids = int(num_sigmoids)
        self.beta = beta
        self.extrema = extrema
        self.minimo = -extrema
        self.massimo = extrema
        self.range_num = torch.arange(self.minimo + 0.5, self.massimo).type(torch.FloatTensor)
        if self.num_sigmoids > 0:
            self.jump = len(self.range_num) / self.num_sigmoids
            self.levels = num_sigmoids + 1
        else:
            self.levels = extrema * 2 + 1
        if self.num_sigmoids == 0:
            self.b = nn.Parameter(self.range_num.type(torch.FloatTensor), requires_grad=trainable)
            self.w = nn.Parameter(torch.ones(len(self.range_num)), requires_grad=trainable)
        else:
            c = len(self.range_num) / self.num_sigmoids
            self.b = nn.Parameter(torch.arange(self.minimo + self.jump / 2, self.massimo + self.jump / 2, c))
            self.w = nn.Parameter(torch.zeros(self.num_sigmoids) + self.jump)
        self.tr_parameters = sum(p.numel() for p in self.parameters() if p.requires_grad)
        self.length = len(self.range_num) if self.num_sigmoids == 0 else self.num_sigmoids
        self.map_sos_cdf = {}
        self.map_cdf_sos = {}
        self.update_state()

    def _all_b(self):
        return self.b

    def _all_w(self):
        return self.w

    def _effective(self, w: Tensor, b: Tensor):
        cum = torch.cat((w.new_zeros(1), torch.cumsum(w, dim=0)))          # the -sum(w)/2 shift cancels in the differences
        return w, torch.sort(b)[0], (cum[1:] - cum[:-1]) / 2

    def update_state(self, device=None):
        device = self.w.device if device is None else device
        before = getattr(self, "_cum_key", None)
        self.update_cumulative_weights(device=device)
        if before == self._cum_key and hasattr(self, "average_points"):
            return                       # nothing changed: the reference calls this every forward
        self.calculate_average_points()
        self.average_points = self.average_points.to(device)
        self.calculate_distance_points()
        self.distance_points = self.distance_points.to(device)
        self.define_channels_map()

    def update_cumulative_weights(self, device=None):
        """activation.py:91-98.  K <= 1024 numbers: summed sequentially on the host so that the levels
        do not depend on the device's scan order (a GPU cumsum differs from the CPU one in the last
        bit); skipped when w has not changed since the last call."""
        device = self.w.device if device is None else device
        key = (id(self.w), self.w.data_ptr(), self.w._version, str(device))       # id / data_ptr: a replaced Parameter restarts _version at 0
        if getattr(self, "_cum_key", None) == key:
            return
        with torch.no_grad():
            w = self.w.detach().cpu()
            n = (torch.sum(w) / 2).item()
            cum_w = torch.zeros(self.length + 1)
            cum_w[0] = 0.0
            cum_w[1:] = torch.cumsum(w, dim=0)
            self.cum_w = torch.sub(cum_w, n).to(device)
        self._cum_key = key

    def reinitialize_weights_and_bias(self):
        if self.num_sigmoids == 0:
            self.w = nn.Parameter(torch.ones(len(self.range_num)))
            self.b = nn.Parameter(self.range_num.type(torch.FloatTensor))
        else:
            self.w = nn.Parameter(torch.zeros(self.num_sigmoids) + self.jump)
            c = len(self.range_num) / self.num_sigmoids
            self.b = nn.Parameter(torch.arange(self.minimo + self.jump / 2, self.massimo + self.jump / 2, c))

    def define_channels_map(self):
        levels = list(self.cum_w.detach().cpu().numpy())
        mapping = list(np.arange(0, len(levels), 1))
        self.map_sos_cdf = dict(zip(levels, mapping))
        self.map_cdf_sos = dict(zip(mapping, levels))

    @property
    def symbol_offset(self) -> int:
        return 0


class SymStanH(_StanHBase):
    """activation.py:157-304."""

    symmetric = True

    def __init__(self, beta, num_sigmoids, extrema=5, trainable=True):
        super().__init__()
        self.num_sigmoids = int(num_sigmoids)
        self.beta = beta
        self.minimo = -extrema
        self.massimo = extrema
        self.range_num = torch.arange(0.5, self.massimo).type(torch.FloatTensor)
        if self.num_sigmoids > 0:
            self.jump = len(self.range_num) / self.num_sigmoids
            self.levels = num_sigmoids + 1
        else:
            self.levels = extrema * 2 + 1
        if self.num_sigmoids == 0:
            self.b = nn.Parameter(self.range_num.type(torch.FloatTensor), requires_grad=trainable)
            self.w = nn.Parameter(torch.ones(len(self.range_num)), requires_grad=trainable)
        else:
            c = len(self.range_num) / self.num_sigmoids
            self.b = nn.Parameter(torch.arange(self.jump / 2, self.massimo + self.jump / 2, c), requires_grad=trainable)
            self.w = nn.Parameter(torch.zeros(self.num_sigmoids) + self.jump, requires_grad=trainable)
        self.tr_parameters = sum(p.numel() for p in self.parameters() if p.requires_grad)
        self.length = len(self.range_num) if self.num_sigmoids == 0 else self.num_sigmoids
        self.map_sos_cdf = {}
        self.map_cdf_sos = {}
        self.update_state()

    def _all_b(self):
        return self.sym_b

    def _all_w(self):
        return self.sym_w

    def _effective(self, w: Tensor, b: Tensor):
        sym_w = torch.cat((torch.flip(w, [0]), w), 0)
        sym_b = torch.cat((torch.flip(-b, [0]), b), 0)
        half = torch.cat((w.new_zeros(1), torch.cumsum(w, dim=0)))
        cum = torch.cat((-torch.flip(half[1:], dims=[0]), half), dim=0)
        return sym_w, torch.sort(sym_b)[0], (cum[1:] - cum[:-1]) / 2

    def update_weights(self):
        self.sym_w = torch.cat((torch.flip(self.w, [0]), self.w), 0)
        self.sym_b = torch.cat((torch.flip(-self.b, [0]), self.b), 0)

    def update_state(self, device=None):
        device = self.w.device if device is None else device
        key = (id(self.w), id(self.b), self.w.data_ptr(), self.b.data_ptr(), self.w._version, self.b._version, str(device))
        if getattr(self, "_cum_key", None) == key and hasattr(self, "average_points"):
            return                       # nothing changed: the reference calls this every forward
        self._cum_key = key
        with torch.no_grad():
            self.update_weights()
            self.update_cumulative_weights()
        self.cum_w = self.cum_w.to(device)
        self.calculate_average_points()
        self.average_points = self.average_points.to(device)
        self.calculate_distance_points()
        self.distance_points = self.distance_points.to(device)

    def update_cumulative_weights(self):
        cum_w = torch.zeros(self.length + 1)
        cum_w[1:] = torch.cumsum(self.w.detach().cpu(), dim=0)      # host-side, see NonSymStanH
        self.cum_w = torch.cat((-torch.flip(cum_w[1:], dims=[0]), cum_w), dim=0)

    def reinitialize_weights_and_bias(self):
        if self.num_sigmoids == 0:
            self.w = nn.Parameter(torch.ones(len(self.range_num)))
            self.b = nn.Parameter(self.range_num.type(torch.FloatTensor))
        else:
            self.w = nn.Parameter(torch.zeros(self.num_sigmoids) + self.jump)
            c = len(self.range_num) / self.num_sigmoids
            self.b = nn.Parameter(torch.arange(self.minimo + self.jump / 2, self.massimo + self.jump / 2, c))

    def define_channels_map(self):
        levels = list(self.cum_w.detach().cpu().numpy())
        minimum = -int(self.cum_w.shape[0]) // 2
        mapping = list(np.arange(minimum, -minimum, 1).astype(int) + 1)
        self.map_sos_cdf = dict(zip(levels, mapping))
        self.map_cdf_sos = dict(zip(mapping, levels))

    def define_channels_map2(self):
        levels = list(self.cum_w.detach().cpu().numpy())
        mapping = list(np.arange(0, len(levels), 1))
        self.map_sos_cdf = dict(zip(levels, mapping))
        self.map_cdf_sos = dict(zip(mapping, levels))

    @property
    def symbol_offset(self) -> int:
        return -(2 * self.length // 2)


def compute_gap(stanh: _StanHBase, inputs: Tensor, beta=None) -> Tensor:
    """tcm_stanh.py:465-478 in one pass over y: |MSE(y, stanh_beta(y)) - MSE(y, stanh_hard(y))|
    (drives the beta annealing, src/training/step.py:46-54).  Returns a 0-d float32 GPU tensor."""
    sums = stanh.gap_sums(inputs, beta)
    n = max(inputs.numel(), 1)
    return torch.abs(sums[0] / n - sums[1] / n).to(torch.float32)


def _stanh_bwd(stanh, inputs, scales, means, training, removing_mean, scale_bound, lik_bound, beta, g_yhat, g_lik,
               want_params: bool = False):
    """reslic_stanh_gc_bwd_f32: (g_y, g_mu, g_sigma, parameter-gradient sums or None).  ``scales`` / ``g_lik`` may be
    None together (quantizer alone)."""
    lib = _cabi.load()
    d = _cabi.new(_cabi.StanhGcBwdDesc)
    keep = []

    def bind(name, t):
        if t is None:
            return
        if t.shape != inputs.shape:
            t = t.expand_as(inputs)
        t, bs, _ = ops.image_major(t.contiguous() if not t.is_contiguous() and name.startswith("g_") else t)
        keep.append(t)
        setattr(d, name, t.data_ptr())
        setattr(d, name + "_bs", bs)

    _, _, n = ops.image_major(inputs)
    bind("y", inputs); bind("mu", means); bind("sigma", scales); bind("g_yhat", g_yhat); bind("g_lik", g_lik)
    B = inputs.shape[0] if inputs.dim() > 0 else 1
    d.B, d.n = B, n
    d.training, d.removing_mean = (1 if training else 0), (1 if removing_mean else 0)
    d.scale_bound = float(scale_bound)
    d.likelihood_bound = float(lik_bound)
    d.tables, tk = stanh._tables(beta)
    keep.append(tk)
    outs = []
    for name in ("g_y", "g_mu", "g_sigma"):
        t = torch.empty(inputs.shape, dtype=torch.float32, device=inputs.device)
        setattr(d, name, t.data_ptr())
        setattr(d, name + "_bs", n)
        outs.append(t)
    g_par = None
    if want_params:
        g_par = torch.empty(5 * int(d.tables.K) + 2, dtype=torch.float64, device=inputs.device)
        d.g_params, d.g_params_len = g_par.data_ptr(), g_par.numel()
    with torch.cuda.device(inputs.device):
        code = lib.reslic_stanh_gc_bwd_f32(C.byref(d), _cabi.current_stream_ptr(inputs.device))
    _cabi.check(code, "reslic_stanh_gc_bwd_f32")
    return (*outs, g_par)


def _stanh_param_grads(stanh, g_par: Tensor, w: Tensor, b: Tensor):
    """Kernel sums (A | Bq | Ww | Wb | Hd, see reslic_stanh_gc_bwd_desc) -> gradients of the raw w / b."""
    K = (g_par.numel() - 2) // 5
    A, Bq, Ww, Wb, Hd = torch.split(g_par, [K + 1, K + 1, K, K, K])
    # dLoss/dw_eff[k] = (sum_{m>k} A[m] - sum_{m<=k} Bq[m]) / 2 + Ww[k]
    g_weff = 0.5 * ((A.sum() - torch.cumsum(A, 0)[:K]) - torch.cumsum(Bq, 0)[:K]) + Ww
    with torch.enable_grad():
        wl = w.detach().double().requires_grad_(True)
        bl = b.detach().double().requires_grad_(True)
        w_eff, b_eff, dist = stanh._effective(wl, bl)
        g_w, g_b = torch.autograd.grad([w_eff, b_eff, dist], [wl, bl], [g_weff, Wb, Hd], allow_unused=True)
    g_w = torch.zeros_like(w) if g_w is None else g_w.to(w.dtype)
    g_b = torch.zeros_like(b) if g_b is None else g_b.to(b.dtype)
    return g_w, g_b


class _StanhQuantizeFn(torch.autograd.Function):
    """The STanH quantizer ALONE as an autograd node: ``stanh_beta(x - mu*[removing_mean]) + mu*[removing_mean]``
    (soft, ``training``) or ``stanh_hard(x - mu) + mu`` — the value of ``quantize(x, "training" | "dequantize", mu)``
    and of the bare activation.  The reference runs these with grad enabled (``quantize(y, mode="training")`` in
    src/models/stanh/tcm_stanh.py:448, wacnn_stanh.py:319; the activation inside update_state); backward is the
    quantizer part of reslic_stanh_gc_bwd_f32 (no likelihood: sigma and g_lik absent), incl. the gradients of the
    trainable STanH parameters."""

    @staticmethod
    def forward(ctx, stanh, inputs, means, training, removing_mean, beta, w, b):
        with torch.no_grad():
            x = inputs.detach()
            m = None if means is None else means.detach()
            if m is not None and m.shape != x.shape:
                m = m.expand_as(x)
            r = _stanh_quantize_launch(stanh, x, m, training, removing_mean, beta)
        ctx.stanh, ctx.cfg = stanh, (bool(training), bool(removing_mean), float(beta))
        ctx.save_for_backward(inputs, means, w, b)
        return r

    @staticmethod
    def backward(ctx, g):
        inputs, means, w, b = ctx.saved_tensors
        training, removing_mean, beta = ctx.cfg
        want_par = ctx.needs_input_grad[6] or ctx.needs_input_grad[7]
        m = means
        if m is not None and m.shape != inputs.shape:
            m = m.expand_as(inputs)
        g_y, g_mu, _, g_par = _stanh_bwd(ctx.stanh, inputs.detach(), None, None if m is None else m.detach(), training,
                                         removing_mean, 0.11, 0.0, beta, g.contiguous(), None, want_params=want_par)
        g_w = g_b = None
        if want_par:
            g_w, g_b = _stanh_param_grads(ctx.stanh, g_par, w, b)
        if means is not None and g_mu.shape != means.shape:       # broadcast means: reduce the gradient back
            g_mu = g_mu.sum_to_size(means.shape)
        return (None, g_y if ctx.needs_input_grad[1] else None, g_mu if (means is not None and ctx.needs_input_grad[2]) else None,
                None, None, None, g_w if ctx.needs_input_grad[6] else None, g_b if ctx.needs_input_grad[7] else None)


def _stanh_quantize_launch(stanh, x: Tensor, means: Optional[Tensor], training: bool, removing_mean: bool, beta: float) -> Tensor:
    """reslic_stanh_gc_fwd_f32 with the quantizer output only (no scales, no likelihood)."""
    lib = _cabi.load()
    ops._require_cuda("inputs", x)
    d = _cabi.new(_cabi.StanhGcDesc)
    keep = []

    def bind(name, t):
        t, bs, n = ops.image_major(t)
        keep.append(t)
        setattr(d, name, t.data_ptr())
        setattr(d, name + "_bs", bs)
        return n

    n = bind("y", x)
    if means is not None:
        ops._require_cuda("means", means)
        bind("mu", means)
    d.B, d.n = (x.shape[0] if x.dim() > 0 else 1), n
    d.training, d.removing_mean = (1 if training else 0), (1 if removing_mean else 0)
    d.scale_bound, d.likelihood_bound = 0.11, 0.0
    d.tables, tk = stanh._tables(beta)
    keep.append(tk)
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    d.yhat, d.yhat_bs = out.data_ptr(), n
    with torch.cuda.device(x.device):
        code = lib.reslic_stanh_gc_fwd_f32(C.byref(d), _cabi.current_stream_ptr(x.device))
    _cabi.check(code, "reslic_stanh_gc_fwd_f32")
    return out


class _EbStanhFn(torch.autograd.Function):
    """Autograd node of EntropyBottleneckStanh.forward: the fused forward (which also leaves every element's bin
    half-widths and level cell), and for the backward two kernels — reslic_eb_bwd_f32 with variable bins (gradient of the
    bounded likelihood w.r.t. the quantizer output, the 58 parameters of every channel and the level gaps) and the STanH
    quantizer's own backward (reslic_stanh_gc_bwd_f32 without a likelihood: d/dz, and d/dw, d/db of a trainable STanH,
    to which the level-gap sums are added as the distance_points term)."""

    @staticmethod
    def forward(ctx, module, x, training, w, b, *params):
        with torch.no_grad():
            r = module._fused(x.detach(), training, ("zhat", "lik", "half_lo", "half_up", "cell"))
        ctx.module, ctx.training = module, bool(training)
        ctx.save_for_backward(x, r["zhat"], r["half_lo"], r["half_up"], r["cell"], w, b, *params)
        ctx.set_materialize_grads(False)
        return r["zhat"], r["lik"]

    @staticmethod
    def backward(ctx, g_zhat, g_lik):
        x, zhat, lo, up, cell, w, b, *params = ctx.saved_tensors
        mod = ctx.module
        m, bi, f = params[:5], params[5:10], params[10:14]
        want_par = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        need_eb = any(ctx.needs_input_grad[5:])
        K = int(mod.stanh.distance_points.numel())
        g_dist = torch.zeros(K, dtype=torch.float64, device=x.device) if want_par else None
        zero_med = torch.zeros(x.shape[1], dtype=torch.float32, device=x.device)
        c = lambda t: None if t is None else t.contiguous()
        g_q, gm, gb, gf, _ = ops.eb_backward(
            zhat, m, bi, f, zero_med, training=True, g_zhat=c(g_zhat), g_lik=c(g_lik), need_z=True, need_params=need_eb,
            likelihood_bound=mod._likelihood_bound if mod.use_likelihood_bound else 0.0,
            half_lo=lo, half_up=up, cell=cell, g_dist=g_dist)
        # through the quantizer z_hat = stanh(z): soft in training (derivative inside the saturation window), hard otherwise (zero)
        g_x, _, _, g_par = _stanh_bwd(mod.stanh, x.detach(), None, None, ctx.training, False, 0.11, 0.0, mod.stanh.beta, g_q, None,
                                      want_params=want_par)
        g_w = g_b = None
        if want_par:
            Kt = (g_par.numel() - 2) // 5
            g_par[-Kt:] += g_dist[:Kt]              # Hd: dLoss / d distance_points (the bins' half-widths)
            g_w, g_b = _stanh_param_grads(mod.stanh, g_par, w, b)
        grads = [None, g_x if ctx.needs_input_grad[1] else None, None,
                 g_w if ctx.needs_input_grad[3] else None, g_b if ctx.needs_input_grad[4] else None]
        if need_eb:
            grads += [*gm, *gb, *gf]
        else:
            grads += [None] * 14
        return tuple(grads)


# ----------------------------------------------------------------------------- entropy model
class _StanhGcFn(torch.autograd.Function):
    """Autograd node around the fused STanH forward / backward kernels.  ``w`` and ``b`` are the module's
    parameters: when either requires grad the backward kernel also accumulates the parameter-gradient
    sums (reslic_stanh_gc_bwd_desc.g_params), which are mapped to w / b here with K-sized torch ops."""

    @staticmethod
    def forward(ctx, module, inputs, scales, means, training, w, b):
        with torch.no_grad():
            r = module._stanh_fused(inputs, scales, means, training, ("yhat", "lik"))
        ctx.module, ctx.training = module, bool(training)
        ctx.save_for_backward(inputs, scales, means, w, b)
        ctx.set_materialize_grads(False)
        return r["yhat"], r["lik"]

    @staticmethod
    def backward(ctx, g_yhat, g_lik):
        inputs, scales, means, w, b = ctx.saved_tensors
        want_par = ctx.needs_input_grad[5] or ctx.needs_input_grad[6]
        g_y, g_mu, g_sigma, g_par = ctx.module._stanh_backward(inputs, scales, means, ctx.training, g_yhat, g_lik,
                                                              want_params=want_par)
        g_w = g_b = None
        if want_par:
            g_w, g_b = ctx.module._stanh_param_grads(g_par, w, b)
            if not ctx.needs_input_grad[5]:
                g_w = None
            if not ctx.needs_input_grad[6]:
                g_b = None
        return None, g_y, g_sigma, (g_mu if means is not None else None), None, g_w, g_b


class HypeEntropyModelSoS(EntropyModel):
    """adaptive_gaussian_conditional.py:17-300: EntropyModel whose quantizer is ``self.stanh``."""

    def __init__(self, removing_mean=True, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None,
                 entropy_coder_precision: int = 16):
        super().__init__(likelihood_bound=likelihood_bound, entropy_coder=entropy_coder,
                         entropy_coder_precision=entropy_coder_precision)
        self.removing_mean = removing_mean

    def define_permutation(self, x):
        perm = np.arange(len(x.shape))
        perm[0], perm[1] = perm[1], perm[0]
        inv_perm = np.arange(len(x.shape))[np.argsort(perm)]
        return perm, inv_perm

    def _stanh_fused(self, inputs: Tensor, scales: Optional[Tensor], means: Optional[Tensor], training: bool,
                     want, beta=None, out: Optional[dict] = None, likelihood_bound: Optional[float] = None):
        lib = _cabi.load()
        ops._require_cuda("inputs", inputs)
        _no_grad_path(inputs, scales, means, self.stanh.w, self.stanh.b)
        want = set(want)
        d = _cabi.new(_cabi.StanhGcDesc)
        keep = []

        def bind(name, t):
            t, bs, n = ops.image_major(t)
            keep.append(t)
            setattr(d, name, t.data_ptr())
            setattr(d, name + "_bs", bs)
            return n

        n = bind("y", inputs)
        if means is not None:
            ops._require_cuda("means", means)
            bind("mu", means.expand_as(inputs) if means.shape != inputs.shape else means)
        if want & {"lik", "bits"}:
            if scales is None:
                raise ValueError("scales are required for the likelihood")
            ops._require_cuda("scales", scales)
            if scales.shape != inputs.shape:
                raise ValueError("scales shape must match inputs")
            bind("sigma", scales)
        B = inputs.shape[0] if inputs.dim() > 0 else 1
        d.B, d.n = B, n
        d.training = int(training) if training in (0, 1, 2) else (1 if training else 0)
        d.removing_mean = 1 if self.removing_mean else 0
        d.scale_bound = float(getattr(self, "_scale_bound", 0.11))
        if likelihood_bound is None:
            likelihood_bound = self._likelihood_bound if self.use_likelihood_bound else 0.0
        d.likelihood_bound = float(likelihood_bound)
        d.tables, tk = self.stanh._tables(self.stanh.beta if beta is None else beta)
        keep.append(tk)
        res = {}
        for name, dtype in (("yhat", torch.float32), ("lik", torch.float32), ("sym", torch.int32), ("ste", torch.float32)):
            if name in want:
                t = (out or {}).get(name)          # caller's buffer (static buffers of a captured graph), else a fresh one
                if t is None:
                    t = torch.empty(inputs.shape, dtype=dtype, device=inputs.device)
                elif t.shape != inputs.shape or t.dtype != dtype or t.device != inputs.device:
                    raise ValueError(f"out['{name}'] must be a {dtype} tensor shaped like the inputs on their device")
                tv, bs, _ = ops.image_major(t)
                if tv.data_ptr() != t.data_ptr():
                    raise ValueError(f"out['{name}'] must be image-major (contiguous per image)")
                setattr(d, name, t.data_ptr())
                setattr(d, name + "_bs", bs)
                res[name] = t
        if "bits" in want:      # out: the rate keys of ops.gc_forward (bits, bits_accumulate, bits_deferred, bits_collect, workspace)
            res["bits"] = ops._rate_outputs(d, dict(out or {}), B, inputs.device, keep)
        with torch.cuda.device(inputs.device):
            code = lib.reslic_stanh_gc_fwd_f32(C.byref(d), _cabi.current_stream_ptr(inputs.device))
        _cabi.check(code, "reslic_stanh_gc_fwd_f32")
        return res

    def _stanh_backward(self, inputs, scales, means, training, g_yhat, g_lik, want_params: bool = False):
        return _stanh_bwd(self.stanh, inputs, scales, means, training, bool(self.removing_mean),
                          float(getattr(self, "_scale_bound", 0.11)),
                          self._likelihood_bound if self.use_likelihood_bound else 0.0, self.stanh.beta, g_yhat, g_lik,
                          want_params)

    def _stanh_param_grads(self, g_par: Tensor, w: Tensor, b: Tensor):
        return _stanh_param_grads(self.stanh, g_par, w, b)

    def quantize(self, inputs, mode, means=None, perms=None):
        """modes "training" | "dequantize" | "symbols" (:95-157).  ``perms`` is accepted for API
        compatibility; the op is elementwise so no permutation is needed."""
        if mode in ("training", "dequantize"):
            training = mode == "training"
            if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (inputs, means, self.stanh.w, self.stanh.b)):
                # the reference runs this with grad enabled (tcm_stanh.py:448, wacnn_stanh.py:319)
                return _StanhQuantizeFn.apply(self.stanh, inputs, means, training, bool(self.removing_mean),
                                              float(self.stanh.beta), self.stanh.w, self.stanh.b)
            return self._stanh_fused(inputs, None, means, training, ("yhat",))["yhat"]
        assert mode == "symbols", mode
        with torch.no_grad():           # integers: nothing to differentiate
            return self._stanh_fused(inputs.detach(), None, None if means is None else means.detach(), False, ("sym",))["sym"]

    def dequantize(self, inputs, means=None, dtype=torch.float):
        """Level index -> level value (+ means) (:174-193), as one gather instead of a Python loop."""
        levels = self.stanh.cum_w.to(inputs.device)
        idx = (inputs.long() - self.stanh.symbol_offset).clamp_(0, levels.numel() - 1)
        outputs = levels[idx].to(dtype)
        if means is not None:
            outputs = outputs.type_as(means) + means
        return outputs.type(dtype)

    def compress(self, symbols, indexes):
        """:268-300 — takes SYMBOLS (the caller quantizes first)."""
        if len(symbols.size()) < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if symbols.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_cdf_size()
        self._check_cdf_length()
        from .entropy_models import _rans

        return _rans().encode_with_indexes_batch(symbols.int(), indexes, self._quantized_cdf,
                                              self._cdf_length.reshape(-1), self._offset.reshape(-1))


class GaussianConditionalStanh(HypeEntropyModelSoS):
    """adaptive_gaussian_conditional.py:312-725."""

    def __init__(self, scale_table: Optional[Union[List, Tuple]], *args: Any, scale_bound: float = 0.11,
                 tail_mass: float = 1e-9, gaussian_configuration=None, channels: int = 128, **kwargs: Any):
        super().__init__(removing_mean=gaussian_configuration["removing_mean"], *args, **kwargs)
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
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)
        self.channels = int(channels)
        self.M = int(channels)
        self.num_sigmoids = int(gaussian_configuration["num_sigmoids"])
        self.extrema = gaussian_configuration["extrema"]
        self.symmetry = gaussian_configuration["symmetry"]
        cls = SymStanH if self.symmetry else NonSymStanH
        self.stanh = cls(gaussian_configuration["beta"], self.num_sigmoids, extrema=self.extrema,
                         trainable=gaussian_configuration["trainable"])

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_cumulative(self, inputs):
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table):
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update(device)
        return True

    def define_v0_and_v1(self, inputs, average_points, distance_points):
        """Setup-time (update()) restatement of :495-537: half-widths of each value's level cell."""
        shape = inputs.shape
        v = inputs.reshape(-1)
        avg = average_points.to(v.device)
        dist = distance_points.to(v.device)
        j = torch.searchsorted(avg.contiguous(), v.contiguous(), right=False)      # #{k : avg_k < v}
        inside = (v > -1000) & (v <= 1000)
        zero = torch.zeros(1, device=v.device, dtype=dist.dtype)
        left = torch.cat((zero, dist))
        right = torch.cat((dist, zero))
        v0 = torch.where(inside, left[j], torch.zeros_like(v)).reshape(shape)
        v1 = torch.where(inside, right[j], torch.zeros_like(v)).reshape(shape)
        return v0, v1

    def update(self, device=None):
        """CDF tables over the STanH levels (:397-454); setup code, once per model."""
        device = self.scale_table.device if device is None else device
        self.stanh.update_state(device)
        max_length = self.stanh.cum_w.shape[0]
        pmf_length = (torch.zeros(self.scale_table.shape[0]).int().to(device) + max_length)
        self.stanh.define_channels_map()
        samples = self.stanh.cum_w.repeat(self.scale_table.shape[0], 1).to(device).float()
        # symbol - offset indexes the CDF row: level indexes start at stanh.symbol_offset
        # (the reference stores -cum_w[0] here, which only matches for integer-spaced levels)
        self._offset = torch.full((self.scale_table.shape[0],), self.stanh.symbol_offset, dtype=torch.int32,
                                  device=device)
        low, up = self.define_v0_and_v1(samples, self.stanh.average_points, self.stanh.distance_points)
        samples_scale = self.scale_table.unsqueeze(1).float()
        upper_pos = self._standardized_cumulative((low - samples) / samples_scale) * (samples >= 0)
        upper_neg = self._standardized_cumulative((samples + up) / samples_scale) * (samples < 0)
        lower_pos = self._standardized_cumulative((-up - samples) / samples_scale) * (samples >= 0)
        lower_neg = self._standardized_cumulative((samples - low) / samples_scale) * (samples < 0)
        upper = upper_pos + upper_neg
        lower = lower_pos + lower_neg
        pmf = upper - lower
        self.pmf = pmf
        self.cdf = self.pmf_to_cdf()
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._cdf_length = (pmf_length + 2).int()

    def pmf_to_cdf(self):
        cdf = self.pmf.cumsum(dim=-1)
        zeros = torch.zeros(self.pmf.shape[:-1] + (1,), dtype=self.pmf.dtype, device=self.pmf.device)
        return torch.cat([zeros, cdf], dim=-1).clamp(max=1.0)

    # ---- per-element path: fused kernel -------------------------------------------------
    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None):
        """Unbounded variable-bin likelihood of already-quantised ``inputs`` (:541-580): the fused
        pass with the quantizer switched off (y_hat = inputs)."""
        return self._stanh_fused(inputs, scales, means, 2, ("lik",), likelihood_bound=0.0)["lik"]

    def forward(self, values, scales, training=True, means=None):
        """:588-603 — note the reference's argument order and default ``training=True``."""
        if training is None:
            training = self.training
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad
                                           for t in (values, scales, means, self.stanh.w, self.stanh.b)):
            return _StanhGcFn.apply(self, values, scales, means, bool(training), self.stanh.w, self.stanh.b)
        r = self._stanh_fused(values, scales, means, bool(training), ("yhat", "lik"))
        return r["yhat"], r["lik"]

    def forward_fused(self, values, scales, training=True, means=None, want=("yhat", "lik", "bits"), out=None):
        return self._stanh_fused(values, scales, means, bool(training), want, out=out)

    def build_indexes(self, scales: Tensor):
        if self.scale_table.numel() == 0:
            raise ValueError("Uninitialized scale_table. Run update_scale_table() first")
        return ops.build_indexes(scales, self.scale_table, self._scale_bound)

    def permutation_function(self, x):
        return self.define_permutation(x)

    def compress(self, x, indexes, perms=None, means=None):
        x = self.quantize(x, "symbols", means=means, perms=perms)
        return super().compress(x, indexes)

    def decompress(self, strings, indexes, means=None, flag=1):
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, len(indexes.size())):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        from .entropy_models import _rans

        symbols = _rans().decode_with_indexes_batch(strings, indexes, self._quantized_cdf,
                                                 self._cdf_length.reshape(-1), self._offset.reshape(-1))
        return self.dequantize(symbols, means=means)


class EntropyBottleneckStanh(EntropyModel):
    """src/entropy_models/adaptive_entropy_bottleneck.py:299-771 (base EntropyModelSoS :24-296): the
    factorized bottleneck whose quantizer is a STanH (no medians) and whose bins follow the STanH
    levels.  forward() is one fused kernel launch; update() keeps the reference's float pmf/cdf.
    The reference's compress()/decompress() for this class reference undefined names (SURVEY.md
    App. B) and are not reproduced."""

    def __init__(self, channels: int, *args: Any, tail_mass: float = 1e-9, pretrained_entropy_model=None,
                 factorized_configuration=None, init_scale: float = 10, filters: Tuple[int, ...] = (3, 3, 3, 3),
                 **kwargs: Any):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.M = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        self.pmf_length = None
        self.num_sigmoids = int(factorized_configuration["num_sigmoids"])
        self.extrema = factorized_configuration["extrema"]
        self.symmetry = factorized_configuration["symmetry"]
        cls = SymStanH if self.symmetry else NonSymStanH
        self.stanh = cls(factorized_configuration["beta"], self.num_sigmoids, extrema=self.extrema,
                         trainable=factorized_configuration["trainable"])
        filt = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            if pretrained_entropy_model is None:
                init = np.log(np.expm1(1 / scale / filt[i + 1]))
                matrix = torch.Tensor(self.channels, filt[i + 1], filt[i])
                matrix.data.fill_(init)
                bias = torch.Tensor(self.channels, filt[i + 1], 1)
                nn.init.uniform_(bias, -0.5, 0.5)
                factor = torch.zeros(self.channels, filt[i + 1], 1) if i < len(self.filters) else None
            else:
                matrix = getattr(pretrained_entropy_model, f"_matrix{i:d}").data.clone()
                bias = getattr(pretrained_entropy_model, f"_bias{i:d}").data.clone()
                factor = getattr(pretrained_entropy_model, f"_factor{i:d}").data.clone() if i < len(self.filters) else None
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if factor is not None:
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        target = np.log(2 / 1e-9 - 1)
        self.target = torch.Tensor([-target, 0, target])

    def _params(self):
        n = len(self.filters) + 1
        return ([getattr(self, f"_matrix{i:d}") for i in range(n)], [getattr(self, f"_bias{i:d}") for i in range(n)],
                [getattr(self, f"_factor{i:d}") for i in range(n - 1)])

    def define_permutation(self, x):
        perm = np.arange(len(x.shape))
        perm[0], perm[1] = perm[1], perm[0]
        inv_perm = np.arange(len(x.shape))[np.argsort(perm)]
        return perm, inv_perm

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        """Setup-time evaluation (update()); the per-element path is the fused kernel (:525-543)."""
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                matrix, bias = matrix.detach(), bias.detach()
            logits = torch.matmul(torch.nn.functional.softplus(matrix), logits) + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def _fused(self, x: Tensor, training: bool, want, beta=None, out: Optional[dict] = None):
        lib = _cabi.load()
        ops._require_cuda("x", x)
        if self.filters != (3, 3, 3, 3):
            raise _cabi.ReslicError("the CUDA bottleneck supports filters=(3,3,3,3) only")
        m, b, f = self._params()
        if torch.is_grad_enabled() and any(t.requires_grad for t in (x, *m, *b, *f, self.stanh.w, self.stanh.b)):
            raise _cabi.ReslicError("_fused / forward_fused record no autograd graph: call under torch.no_grad(), or use forward()")
        xc = x.contiguous()
        B, Cc = xc.shape[0], xc.shape[1]
        hw = 1
        for s_ in xc.shape[2:]:
            hw *= s_
        d = _cabi.new(_cabi.EbStanhDesc)
        keep = [xc]
        d.z, d.z_bs = xc.data_ptr(), Cc * hw
        d.B, d.C, d.hw = B, Cc, hw
        d.training = 1 if training else 0
        d.likelihood_bound = self._likelihood_bound if self.use_likelihood_bound else 0.0
        for i in range(5):
            mi, bi = m[i].detach().contiguous(), b[i].detach().contiguous()
            keep += [mi, bi]
            d.matrix[i], d.bias[i] = mi.data_ptr(), bi.data_ptr()
        for i in range(4):
            fi = f[i].detach().contiguous()
            keep.append(fi)
            d.factor[i] = fi.data_ptr()
        d.tables, tk = self.stanh._tables(self.stanh.beta if beta is None else beta)
        keep.append(tk)
        res = {}
        for name, dtype in (("zhat", torch.float32), ("lik", torch.float32), ("sym", torch.int32),
                            ("half_lo", torch.float32), ("half_up", torch.float32), ("cell", torch.int32)):
            if name in want:
                t = torch.empty(xc.shape, dtype=dtype, device=x.device)
                setattr(d, name, t.data_ptr())
                setattr(d, name + "_bs", Cc * hw)
                res[name] = t
        if "bits" in want:
            res["bits"] = ops._rate_outputs(d, dict(out or {}), B, x.device, keep)
        with torch.cuda.device(x.device):
            code = lib.reslic_eb_stanh_fwd_f32(C.byref(d), _cabi.current_stream_ptr(x.device))
        _cabi.check(code, "reslic_eb_stanh_fwd_f32")
        return res

    def quantize(self, inputs, mode, means=None, perms=None):
        """EntropyModelSoS.quantize (:113-177): "training" (soft STanH, means ignored as in the
        reference), "dequantize" (hard levels about the means), "symbols" (level index)."""
        grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (inputs, means, self.stanh.w, self.stanh.b))
        if mode == "training":
            if grad:    # the quantizer alone: differentiable through the STanH backward kernel
                return _StanhQuantizeFn.apply(self.stanh, inputs, None, True, False, float(self.stanh.beta), self.stanh.w, self.stanh.b)
            with torch.no_grad():
                return self._fused(inputs, True, ("zhat",))["zhat"]
        if mode == "dequantize":
            if grad:
                return _StanhQuantizeFn.apply(self.stanh, inputs, means, False, True, float(self.stanh.beta), self.stanh.w, self.stanh.b)
            with torch.no_grad():
                x = inputs - means if means is not None else inputs
                out = self._fused(x, False, ("zhat",))["zhat"]
                return out + means if means is not None else out
        assert mode == "symbols", mode
        with torch.no_grad():
            x = inputs.detach() - means.detach() if means is not None else inputs.detach()
            return self._fused(x, False, ("sym",))["sym"]

    def forward(self, x: Tensor, training: bool = True):
        """:679-708 — note the reference default ``training=True``."""
        m, b, f = self._params()
        if torch.is_grad_enabled() and any(t.requires_grad for t in (x, *m, *b, *f, self.stanh.w, self.stanh.b)):
            # inside a training step (src/models/stanh/wacnn_stanh.py:160, balle18_stanh.py:26,124)
            return _EbStanhFn.apply(self, x, bool(training), self.stanh.w, self.stanh.b, *m, *b, *f)
        r = self._fused(x, bool(training), ("zhat", "lik"))
        return r["zhat"], r["lik"]

    def forward_fused(self, x: Tensor, training: bool = True, want=("zhat", "lik", "bits"), out=None):
        return self._fused(x, bool(training), want, out=out)

    def update(self, device=None):
        """Float pmf / cdf over the STanH levels per channel (:481-514)."""
        device = self._matrix0.device if device is None else device
        self.stanh.update_state(device)
        with torch.no_grad():
            samples = self.stanh.cum_w.repeat(self.M, 1).unsqueeze(1).to(device)
            avg, dist = self.stanh.average_points.to(device), self.stanh.distance_points.to(device)
            flat = samples.reshape(-1)
            j = torch.searchsorted(avg.contiguous(), flat.contiguous(), right=False)
            inside = (flat > -1000) & (flat <= 1000)
            zero = torch.zeros(1, device=device, dtype=dist.dtype)
            low = torch.where(inside, torch.cat((zero, dist))[j], torch.zeros_like(flat))
            up = torch.where(inside, torch.cat((dist, zero))[j], torch.zeros_like(flat))
            v0 = (flat - low).reshape(samples.shape)
            v1 = (flat + up).reshape(samples.shape)
            lower = self._logits_cumulative(v0, stop_gradient=True)
            upper = self._logits_cumulative(v1, stop_gradient=True)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
            self.pmf = pmf
            cdf = pmf.cumsum(dim=-1)
            self.cdf = torch.cat([torch.zeros(pmf.shape[:-1] + (1,), dtype=pmf.dtype, device=device), cdf],
                                 dim=-1).clamp(max=1.0)
        return True

    def order_pars(self):
        self.stanh.w = nn.Parameter(torch.sort(self.stanh.w)[0])
        self.stanh.b = nn.Parameter(torch.sort(self.stanh.b)[0])
