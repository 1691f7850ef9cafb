"""Functional entry points: torch tensors in, one fused CUDA launch, torch tensors out.

These are thin marshalling layers over the C ABI (``include/reslic_b200.h``); all
arithmetic happens in the sm_100a kernels.  CPU tensors are rejected — there is no
fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _cabi


@dataclass
class GcOutputs:
    yhat: Optional[Tensor] = None     # quantize() output (noisy / rounded)
    ste: Optional[Tensor] = None      # round(y - mu) + mu
    lik: Optional[Tensor] = None      # bounded likelihood
    sym: Optional[Tensor] = None      # int32 symbols
    idx: Optional[Tensor] = None      # int32 scale-table indexes
    bits: Optional[Tensor] = None     # [B] float64, -sum log2 L per image


@dataclass
class EbOutputs:
    zhat: Optional[Tensor] = None
    ste: Optional[Tensor] = None
    lik: Optional[Tensor] = None
    sym: Optional[Tensor] = None
    bits: Optional[Tensor] = None


def _require_cuda(name: str, t: Tensor, dtype=torch.float32):
    if not isinstance(t, Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise _cabi.ReslicError(
            f"{name} is on {t.device}: the entropy-model path runs only on CUDA (sm_100a); there is no CPU fallback"
        )
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def image_major(t: Tensor) -> Tuple[Tensor, int, int]:
    """Return (tensor, batch_stride, elems_per_image) such that image b is the contiguous
    run [b*batch_stride, b*batch_stride + n).  A channel slice of a contiguous NCHW tensor
    (``y.chunk(5, 1)[k]``, tcm.py:438) qualifies without a copy."""
    if t.dim() < 1:
        t = t.reshape(1)
    B = t.shape[0]
    n = 1
    for s in t.shape[1:]:
        n *= s
    ok = True
    expect = 1
    for size, stride in zip(reversed(t.shape[1:]), reversed(t.stride()[1:])):
        if size != 1 and stride != expect:
            ok = False
            break
        expect *= size
    if not ok:
        t = t.contiguous()
    bs = t.stride(0) if (B > 1 and t.dim() > 0) else n
    return t, int(bs), int(n)


def _out_like(x: Tensor, dtype) -> Tensor:
    return torch.empty(x.shape, dtype=dtype, device=x.device)


def _rate_outputs(d, out: dict, B: int, device, keep: list) -> Optional[Tensor]:
    """Fill the rate fields of a *_fwd descriptor.  ``out["bits_deferred"]`` selects
    RESLIC_RATE_DEFERRED: the sum stays in ``out["workspace"]`` (required then) until
    :func:`rate_finalize` or a later launch with ``out["bits_collect"]`` (RESLIC_RATE_COLLECT: its
    bits = its own sum + everything deferred); otherwise ``out["bits"]`` (or a fresh tensor) receives
    this launch's sum."""
    ws = out.get("workspace")
    if out.get("bits_deferred"):
        if ws is None:
            raise ValueError("out['bits_deferred'] needs the caller's out['workspace'] (it carries the sum)")
        d.bits = None
        d.bits_accumulate = _cabi.RATE_DEFERRED
        bits = None
    else:
        bits = out.get("bits")
        if bits is None:
            bits = torch.empty(B, dtype=torch.float64, device=device)
        elif bits.shape != (B,) or bits.dtype != torch.float64 or not bits.is_contiguous():
            raise ValueError("out['bits'] must be a contiguous float64 [B] tensor")
        d.bits = bits.data_ptr()
        d.bits_accumulate = _cabi.RATE_COLLECT if out.get("bits_collect") else (1 if out.get("bits_accumulate") else 0)
    if ws is None:
        ws = _cabi.workspace(device, B)
    d.workspace = ws.data_ptr()
    d.workspace_bytes = ws.numel()
    keep.append(ws)
    return bits


def rate_finalize(workspace: Tensor, B: int, bits: Optional[Tensor] = None, accumulate: bool = False) -> Tensor:
    """bits[b] (= or +=) the rate the deferred launches left in ``workspace`` (reslic_rate_finalize_f64)."""
    lib = _cabi.load()
    if bits is None:
        if accumulate:
            raise ValueError("accumulate=True needs the bits tensor to add to")
        bits = torch.empty(B, dtype=torch.float64, device=workspace.device)
    elif bits.shape != (B,) or bits.dtype != torch.float64 or not bits.is_contiguous() or not bits.is_cuda:
        raise ValueError("bits must be a contiguous CUDA float64 [B] tensor")
    with torch.cuda.device(workspace.device):
        code = lib.reslic_rate_finalize_f64(workspace.data_ptr(), workspace.numel(), B, bits.data_ptr(),
                                            1 if accumulate else 0, _cabi.current_stream_ptr(workspace.device))
    _cabi.check(code, "reslic_rate_finalize_f64")
    return bits


_PREFETCH_NEXT = os.environ.get("RESLIC_PREFETCH_NEXT", "1") != "0"


def next_channel_slice(y: Tensor) -> Optional[Tensor]:
    """If ``y`` is a channel-slice view [B, C, ...] of a wider contiguous tensor that still has room for another C
    channels behind it (``y.chunk(5, 1)[k]`` with k < 4, tcm.py:438-443), the view of those next C channels — the
    L2 prefetch hint of :func:`gc_forward`; else None.  Only ever describes memory inside ``y``'s own storage."""
    if y.dim() < 2 or y.dtype != torch.float32:
        return None
    B, C = y.shape[0], y.shape[1]
    inner = 1
    for s_ in y.shape[2:]:
        inner *= s_
    if inner == 0 or C == 0 or y.stride(1) != inner:
        return None
    expect = 1
    for size, stride in zip(reversed(y.shape[2:]), reversed(y.stride()[2:])):
        if size != 1 and stride != expect:
            return None
        expect *= size
    row = y.stride(0)                      # elements per image of the wider tensor (kept by chunk / slicing even for B = 1)
    if row < 2 * C * inner or row % inner:
        return None
    off = y.storage_offset() % row         # position inside the image, assuming the wider tensor starts on a row boundary
    if off % inner or off + 2 * C * inner > row:
        return None
    start = y.storage_offset() + C * inner
    if (start + (B - 1) * row + C * inner) * y.element_size() > y.untyped_storage().nbytes():
        return None
    return y.detach().as_strided(y.shape, y.stride(), start)


def rate_from_likelihood(likelihoods: Tensor, out: Optional[dict] = None) -> Optional[Tensor]:
    """bits[b] = -sum log2 likelihoods[b] (float64 [B]) of an existing likelihood tensor [B, ...] in one read pass
    (reslic_rate_from_likelihood_f32; the reference: torch.log(L).sum() / -ln 2, training/loss.py:22-25).
    ``out``: the rate keys of :func:`gc_forward` (bits, bits_accumulate, bits_deferred, bits_collect, workspace)."""
    lib = _cabi.load()
    _require_cuda("likelihoods", likelihoods)
    t, bs, n = image_major(likelihoods)
    B = likelihoods.shape[0] if likelihoods.dim() > 0 else 1

    class _D:      # the rate fields _rate_outputs fills
        bits = None
        bits_accumulate = 0
        workspace = None
        workspace_bytes = 0

    d, keep = _D(), [t]
    bits = _rate_outputs(d, dict(out or {}), B, likelihoods.device, keep)
    with torch.cuda.device(likelihoods.device):
        code = lib.reslic_rate_from_likelihood_f32(t.data_ptr(), bs, B, n, d.bits, d.bits_accumulate, d.workspace,
                                                   d.workspace_bytes, _cabi.current_stream_ptr(likelihoods.device))
    _cabi.check(code, "reslic_rate_from_likelihood_f32")
    return bits


def gc_forward(
    y: Tensor,
    scales: Optional[Tensor] = None,
    means: Optional[Tensor] = None,
    *,
    training: bool = False,
    noise: Optional[Tensor] = None,
    scale_table: Optional[Tensor] = None,
    scale_bound: float = 0.11,
    likelihood_bound: float = 1e-9,
    want: Sequence[str] = ("yhat", "lik"),
    out: Optional[dict] = None,
    seed: int = 0,
    offset: int = 0,
    next_y: Optional[Tensor] = None,
    exchange=None,
    exchange_step: int = 0,
) -> GcOutputs:
    """One fused pass over a Gaussian-conditional slice.

    ``want`` ⊆ {"yhat","ste","lik","sym","idx","bits"}; ``out`` may map any of those names
    to a preallocated tensor (e.g. a channel slice of a full ``y_hat``) to write into.
    Mirrors GaussianConditional.forward + ste_round + build_indexes + quantize("symbols")
    + the log2-rate sum (tcm.py:455,457,544,548; loss.py:24-27).
    ``exchange``: a :class:`reslic_tcm_b200.dist.PeerRateExchange` — the launch that completes ``bits`` (rate mode 0 or
    collect) also publishes the batch's packed rate row to every rank (no collective kernel) as step
    ``exchange.cursor + exchange_step``; the caller advances the cursor (``exchange.advance``).
    """
    lib = _cabi.load()
    _require_cuda("inputs", y)
    want = set(want)
    bad = want - {"yhat", "ste", "lik", "sym", "idx", "bits"}
    if bad:
        raise ValueError(f"unknown outputs requested: {sorted(bad)}")
    if not want:
        raise ValueError("no output requested")
    if want & {"lik", "bits", "idx"}:
        if scales is None:
            raise ValueError("scales are required for likelihood / indexes")
        _require_cuda("scales", scales)
        if scales.shape != y.shape:
            raise ValueError(f"scales shape {tuple(scales.shape)} != inputs shape {tuple(y.shape)}")
    else:
        scales = None
    if means is not None:
        _require_cuda("means", means)
        if means.shape != y.shape:
            means = means.expand_as(y)
    if noise is not None:
        _require_cuda("noise", noise)
        if noise.shape != y.shape:
            raise ValueError("noise shape must match inputs")
    if "idx" in want:
        if scale_table is None or scale_table.numel() == 0:
            raise ValueError("build_indexes needs a scale_table (call update_scale_table first)")
        _require_cuda("scale_table", scale_table)
        scale_table = scale_table.contiguous()
    out = dict(out or {})
    shape = y.shape
    B = shape[0] if y.dim() > 0 else 1
    d = _cabi.new(_cabi.GcDesc)
    keep = []  # keep temporaries alive until the launch is enqueued

    def bind_in(name, t):
        t, bs, n = image_major(t)
        keep.append(t)
        setattr(d, name, t.data_ptr())
        setattr(d, name + "_bs", bs)
        return n

    n = bind_in("y", y)
    if scales is not None:
        bind_in("sigma", scales)
    if means is not None:
        bind_in("mu", means)
    if noise is not None:
        bind_in("noise", noise)
    d.B, d.n = B, n
    d.mode = _cabi.Q_NOISE if training else _cabi.Q_DEQUANTIZE
    d.scale_bound = float(scale_bound)
    d.likelihood_bound = float(likelihood_bound)
    if "idx" in want:
        d.scale_table = scale_table.data_ptr()
        d.table_len = scale_table.numel()
        keep.append(scale_table)
    res = GcOutputs()
    copies = []
    for name, dtype in (("yhat", torch.float32), ("ste", torch.float32), ("lik", torch.float32),
                        ("sym", torch.int32), ("idx", torch.int32)):
        if name not in want:
            continue
        t = out.get(name)
        if t is None:
            t = _out_like(y, dtype)
        else:
            if t.shape != shape or t.dtype != dtype or not t.is_cuda:
                raise ValueError(f"out[{name!r}] has wrong shape/dtype/device")
        tm, bs, _ = image_major(t)
        if tm is not t:  # caller's buffer is not image-major: compute into a temp, copy after
            copies.append((t, tm))
        keep.append(tm)
        setattr(d, name, tm.data_ptr())
        setattr(d, name + "_bs", bs)
        setattr(res, name, t)
    if "bits" in want:
        res.bits = _rate_outputs(d, out, B, y.device, keep)
    d.philox_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    d.philox_offset = int(offset) & 0xFFFFFFFFFFFFFFFF
    if next_y is None and _PREFETCH_NEXT:
        next_y = next_channel_slice(y)      # TCM's slice loop through the module API: the next chunk of the same latent
    if next_y is not None and next_y.is_cuda and next_y.dtype == torch.float32 and next_y.shape == y.shape:
        # hint only (L2 prefetch of the next launch's y, e.g. TCM's next channel slice): a view that is not
        # image-major is simply not prefetched
        tv, bs_n, _ = image_major(next_y)
        if tv.data_ptr() == next_y.data_ptr():
            d.next_y, d.next_y_bs = next_y.data_ptr(), bs_n
            keep.append(next_y)
    if exchange is not None:
        if exchange.device != y.device:
            raise ValueError("exchange lives on another device")
        exchange.desc.step = int(exchange_step)      # read by the library during the call below
        d.exchange = C.pointer(exchange.desc)
    with torch.cuda.device(y.device):
        code = lib.reslic_gc_fwd_f32(C.byref(d), _cabi.current_stream_ptr(y.device))
    _cabi.check(code, "reslic_gc_fwd_f32")
    for dst, src in copies:
        dst.copy_(src)
    return res


def gc_backward(
    y: Tensor,
    scales: Tensor,
    means: Optional[Tensor],
    *,
    training: bool,
    noise: Optional[Tensor] = None,
    scale_bound: float = 0.11,
    likelihood_bound: float = 1e-9,
    g_yhat: Optional[Tensor] = None,
    g_ste: Optional[Tensor] = None,
    g_lik: Optional[Tensor] = None,
    need: Sequence[bool] = (True, True, True),
    seed: int = 0,
    offset: int = 0,
) -> Tuple[Optional[Tensor], Optional[Tensor], Optional[Tensor]]:
    """Backward of gc_forward: (g_y, g_mu, g_sigma) from the upstream gradients of the quantize
    output, the ste_round output and the bounded likelihood (each may be None = zero)."""
    lib = _cabi.load()
    _require_cuda("inputs", y)
    d = _cabi.new(_cabi.GcBwdDesc)
    keep = []

    def bind(name, t):
        if t is None:
            return
        _require_cuda(name, t)
        if t.shape != y.shape:
            t = t.expand_as(y)
        t, bs, _ = image_major(t)
        keep.append(t)
        setattr(d, name, t.data_ptr())
        setattr(d, name + "_bs", bs)

    _, _, n = image_major(y)
    bind("y", y)
    bind("mu", means)
    bind("sigma", scales)
    bind("noise", noise)
    bind("g_yhat", g_yhat)
    bind("g_ste", g_ste)
    bind("g_lik", g_lik)
    d.B, d.n = (y.shape[0] if y.dim() > 0 else 1), n
    d.mode = _cabi.Q_NOISE if training else _cabi.Q_DEQUANTIZE
    d.scale_bound, d.likelihood_bound = float(scale_bound), float(likelihood_bound)
    outs = []
    for name, want in zip(("g_y", "g_mu", "g_sigma"), need):
        t = torch.empty_like(y, memory_format=torch.contiguous_format) if want else None
        if t is not None:
            setattr(d, name, t.data_ptr())
            setattr(d, name + "_bs", n)
        outs.append(t)
    if not any(o is not None for o in outs):
        return None, None, None
    d.philox_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    d.philox_offset = int(offset) & 0xFFFFFFFFFFFFFFFF
    with torch.cuda.device(y.device):
        code = lib.reslic_gc_bwd_f32(C.byref(d), _cabi.current_stream_ptr(y.device))
    _cabi.check(code, "reslic_gc_bwd_f32")
    return tuple(outs)


def build_indexes(scales: Tensor, scale_table: Tensor, scale_bound: float = 0.11) -> Tensor:
    """adaptive_gaussian_conditional.py:606-617 as one launch (int32, same shape)."""
    lib = _cabi.load()
    _require_cuda("scales", scales)
    _require_cuda("scale_table", scale_table)
    s = scales.contiguous()
    tab = scale_table.contiguous()
    idx = torch.empty(s.shape, dtype=torch.int32, device=s.device)
    with torch.cuda.device(s.device):
        code = lib.reslic_build_indexes_f32(s.data_ptr(), s.numel(), float(scale_bound), tab.data_ptr(),
                                            tab.numel(), idx.data_ptr(), _cabi.current_stream_ptr(s.device))
    _cabi.check(code, "reslic_build_indexes_f32")
    return idx


def rans_slots(symbols: Tensor, indexes: Tensor, cdf: Tensor, cdf_length: Tensor, offset: Tensor,
               esc_capacity: Optional[int] = None, out: Optional[Tuple[Tensor, Tensor, Tensor, Tensor]] = None):
    """Device-side front end of the rANS coder (reslic_rans_slots_u32): the per-symbol table lookup of
    ``compressai.ans.encode_with_indexes`` (tcm.py:522,564-565) for int32 ``symbols`` / ``indexes`` of any shape.
    Returns ``(slots, esc_pos, esc_raw, status)`` — uint32-in-int32 slots shaped like ``symbols``
    (``start << 16 | range``), the unordered escape list and the int32[2] status (escape count, error bits), all on
    the device; :func:`reslic_tcm_b200.rans.encode_slots_batch` turns them into strings.  ``out``: preallocated
    ``(slots, esc_pos, esc_raw, status)`` to write into (static buffers of a pipeline)."""
    lib = _cabi.load()
    for name, t in (("symbols", symbols), ("indexes", indexes)):
        _require_cuda(name, t, torch.int32)
    if symbols.shape != indexes.shape:
        raise ValueError("symbols and indexes must have the same shape")
    dev = symbols.device
    s, i = symbols.contiguous(), indexes.contiguous()
    c = cdf.detach().to(device=dev, dtype=torch.int32).contiguous()
    sizes = cdf_length.detach().reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
    offs = offset.detach().reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
    if c.dim() != 2 or sizes.numel() != c.shape[0] or offs.numel() != c.shape[0]:
        raise ValueError("cdf must be [n_cdfs, stride] with one length and one offset per row")
    n = s.numel()
    if out is not None:
        slots, esc_pos, esc_raw, status = out
        if (slots.numel() != n or slots.dtype != torch.int32 or not slots.is_contiguous() or esc_pos.dtype != torch.int32
                or esc_raw.dtype != torch.int64 or esc_raw.numel() != esc_pos.numel() or status.numel() != 2
                or status.dtype != torch.int32 or any(t.device != dev for t in out)):
            raise ValueError("rans_slots: out buffers have the wrong shape, dtype or device")
        cap = esc_pos.numel()
    else:
        cap = int(esc_capacity) if esc_capacity is not None else n // 64 + 4096
        slots = torch.empty(s.shape, dtype=torch.int32, device=dev)
        esc_pos = torch.empty(cap, dtype=torch.int32, device=dev)
        esc_raw = torch.empty(cap, dtype=torch.int64, device=dev)
        status = torch.empty(2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        code = lib.reslic_rans_slots_u32(s.data_ptr(), i.data_ptr(), n, c.data_ptr(), c.shape[0], c.shape[1],
                                         sizes.data_ptr(), offs.data_ptr(), slots.data_ptr(), esc_pos.data_ptr(),
                                         esc_raw.data_ptr(), cap, status.data_ptr(), _cabi.current_stream_ptr(dev))
    _cabi.check(code, "reslic_rans_slots_u32")
    return slots, esc_pos, esc_raw, status


def dequantize(symbols: Tensor, means: Optional[Tensor] = None) -> Tensor:
    """EntropyModel.dequantize: float(sym) + means (tcm.py:623)."""
    lib = _cabi.load()
    _require_cuda("symbols", symbols, torch.int32)
    s = symbols.contiguous()
    m = None
    if means is not None:
        _require_cuda("means", means)
        m = means.expand_as(s).contiguous()
    out = torch.empty(s.shape, dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        code = lib.reslic_dequantize_f32(s.data_ptr(), _cabi.ptr(m), s.numel(), out.data_ptr(),
                                         _cabi.current_stream_ptr(s.device))
    _cabi.check(code, "reslic_dequantize_f32")
    return out


def _eb_fill_params(d, matrices, biases, factors, medians, Cc: int, keep: list) -> None:
    for i in range(5):
        m = matrices[i].detach().contiguous()
        b = biases[i].detach().contiguous()
        _require_cuda(f"_matrix{i}", m)
        _require_cuda(f"_bias{i}", b)
        keep += [m, b]
        d.matrix[i] = m.data_ptr()
        d.bias[i] = b.data_ptr()
    for i in range(4):
        f = factors[i].detach().contiguous()
        _require_cuda(f"_factor{i}", f)
        keep.append(f)
        d.factor[i] = f.data_ptr()
    med = medians.detach().reshape(-1).contiguous()
    _require_cuda("medians", med)
    if med.numel() != Cc:
        raise ValueError("medians must have one entry per channel")
    keep.append(med)
    d.medians = med.data_ptr()


def eb_build_lut(matrices: Sequence[Tensor], biases: Sequence[Tensor], factors: Sequence[Tensor], medians: Tensor,
                 *, likelihood_bound: float = 1e-9) -> Tensor:
    """[C, 130] eval-mode table of the bottleneck (reslic_eb_build_lut_f32): per channel the bounded
    likelihoods of round(z - median) = -32..32 and their log2, bit-identical to what eb_forward computes
    itself.  Valid until a parameter, the medians or the bound changes."""
    lib = _cabi.load()
    if len(matrices) != 5 or len(biases) != 5 or len(factors) != 4:
        raise _cabi.ReslicError("the CUDA bottleneck supports filters=(3,3,3,3) only")
    Cc = matrices[0].shape[0]
    d = _cabi.new(_cabi.EbDesc)
    keep: list = []
    _eb_fill_params(d, matrices, biases, factors, medians, Cc, keep)
    d.C = Cc
    d.likelihood_bound = float(likelihood_bound)
    dev = matrices[0].device
    lut = torch.empty((Cc, _cabi.EB_LUT_STRIDE), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        code = lib.reslic_eb_build_lut_f32(C.byref(d), lut.data_ptr(), _cabi.current_stream_ptr(dev))
    _cabi.check(code, "reslic_eb_build_lut_f32")
    return lut


def eb_forward(
    z: Tensor,
    matrices: Sequence[Tensor],
    biases: Sequence[Tensor],
    factors: Sequence[Tensor],
    medians: Tensor,
    *,
    training: bool = False,
    noise: Optional[Tensor] = None,
    likelihood_bound: float = 1e-9,
    want: Sequence[str] = ("zhat", "lik"),
    out: Optional[dict] = None,
    seed: int = 0,
    offset: int = 0,
    lut: Optional[Tensor] = None,
    next_y: Optional[Tensor] = None,
) -> EbOutputs:
    """One fused pass of the factorized bottleneck over z [B, C, *spatial] (tcm.py:429-433).
    ``lut``: the table from :func:`eb_build_lut` for these very parameters (eval mode only; ignored
    when ``training``) — takes the table construction off the launch."""
    lib = _cabi.load()
    _require_cuda("z", z)
    if z.dim() < 2:
        raise ValueError("z must be at least [B, C]")
    if len(matrices) != 5 or len(biases) != 5 or len(factors) != 4:
        raise _cabi.ReslicError("the CUDA bottleneck supports filters=(3,3,3,3) only")
    B, Cc = z.shape[0], z.shape[1]
    expect_m = [(Cc, 3, 1), (Cc, 3, 3), (Cc, 3, 3), (Cc, 3, 3), (Cc, 1, 3)]
    for i, m in enumerate(matrices):
        if tuple(m.shape) != expect_m[i]:
            raise _cabi.ReslicError(f"_matrix{i} has shape {tuple(m.shape)}; filters must be (3,3,3,3)")
    want = set(want)
    bad = want - {"zhat", "ste", "lik", "sym", "bits"}
    if bad:
        raise ValueError(f"unknown outputs requested: {sorted(bad)}")
    zc = z.contiguous()
    hw = 1
    for s in z.shape[2:]:
        hw *= s
    d = _cabi.new(_cabi.EbDesc)
    keep = [zc]
    d.z, d.z_bs = zc.data_ptr(), Cc * hw
    if noise is not None:
        _require_cuda("noise", noise)
        nc = noise.contiguous()
        keep.append(nc)
        d.noise, d.noise_bs = nc.data_ptr(), Cc * hw
    d.B, d.C, d.hw = B, Cc, hw
    d.mode = _cabi.Q_NOISE if training else _cabi.Q_DEQUANTIZE
    d.likelihood_bound = float(likelihood_bound)
    _eb_fill_params(d, matrices, biases, factors, medians, Cc, keep)
    if lut is not None and not training:
        _require_cuda("lut", lut)
        if lut.shape != (Cc, _cabi.EB_LUT_STRIDE) or lut.dtype != torch.float32 or not lut.is_contiguous():
            raise ValueError(f"lut must be a contiguous float32 [C, {_cabi.EB_LUT_STRIDE}] tensor from eb_build_lut")
        keep.append(lut)
        d.lut = lut.data_ptr()
    res = EbOutputs()
    out = dict(out or {})
    for name, dtype in (("zhat", torch.float32), ("ste", torch.float32), ("lik", torch.float32),
                        ("sym", torch.int32)):
        if name in want:
            t = out.get(name)
            if t is None:
                t = torch.empty(zc.shape, dtype=dtype, device=z.device)
            elif t.shape != zc.shape or t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
                raise ValueError(f"out[{name!r}] must be a contiguous {dtype} CUDA tensor shaped like z")
            setattr(d, name, t.data_ptr())
            setattr(d, name + "_bs", Cc * hw)
            setattr(res, name, t)
    if "bits" in want:
        res.bits = _rate_outputs(d, out, B, z.device, keep)
    d.philox_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    d.philox_offset = int(offset) & 0xFFFFFFFFFFFFFFFF
    if next_y is not None and next_y.is_cuda and next_y.dtype == torch.float32 and next_y.dim() > 0 and next_y.shape[0] == B:
        tv, bs_n, n_n = image_major(next_y)        # hint only: the first y slice the following launch reads
        if tv.data_ptr() == next_y.data_ptr():
            d.next_y, d.next_y_bs, d.next_y_n = next_y.data_ptr(), bs_n, n_n
            keep.append(next_y)
    with torch.cuda.device(z.device):
        code = lib.reslic_eb_fwd_f32(C.byref(d), _cabi.current_stream_ptr(z.device))
    _cabi.check(code, "reslic_eb_fwd_f32")
    return res


_eb_bwd_ws = {}


def _eb_bwd_workspace(device: torch.device, channels: int) -> Tensor:
    """Zero-initialised scratch of reslic_eb_bwd_f32, one per (device, stream) — launches on one stream are
    ordered, and every launch leaves its arrival counters zeroed."""
    need = int(_cabi.load().reslic_eb_bwd_workspace_bytes(channels))
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _eb_bwd_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _eb_bwd_ws[key] = ws
    return ws


def eb_backward(
    z: Tensor,
    matrices: Sequence[Tensor],
    biases: Sequence[Tensor],
    factors: Sequence[Tensor],
    medians: Tensor,
    *,
    training: bool,
    noise: Optional[Tensor] = None,
    likelihood_bound: float = 1e-9,
    g_zhat: Optional[Tensor] = None,
    g_lik: Optional[Tensor] = None,
    need_z: bool = True,
    need_params: bool = True,
    seed: int = 0,
    offset: int = 0,
    half_lo: Optional[Tensor] = None,
    half_up: Optional[Tensor] = None,
    cell: Optional[Tensor] = None,
    g_dist: Optional[Tensor] = None,
):
    """Backward of eb_forward: (g_z, [g_matrix]*5, [g_bias]*5, [g_factor]*4, g_medians).
    Variable bins (EntropyBottleneckStanh): ``z`` is the STanH quantizer's output, ``half_lo`` / ``half_up`` / ``cell`` what
    the forward wrote, ``medians`` zeros; ``g_z`` is then the gradient w.r.t. the quantizer output and ``g_dist``
    (float64 [K], zeroed by the caller) receives dLoss / d distance_points."""
    lib = _cabi.load()
    _require_cuda("z", z)
    B, Cc = z.shape[0], z.shape[1]
    hw = 1
    for s_ in z.shape[2:]:
        hw *= s_
    zc = z.contiguous()
    d = _cabi.new(_cabi.EbBwdDesc)
    keep = [zc]
    d.z, d.z_bs = zc.data_ptr(), Cc * hw
    for name, t in (("noise", noise), ("g_zhat", g_zhat), ("g_lik", g_lik)):
        if t is not None:
            _require_cuda(name, t)
            tc = t.contiguous()
            keep.append(tc)
            setattr(d, name, tc.data_ptr())
            setattr(d, name + "_bs", Cc * hw)
    d.B, d.C, d.hw = B, Cc, hw
    d.mode = _cabi.Q_NOISE if training else _cabi.Q_DEQUANTIZE
    if half_lo is not None or half_up is not None or cell is not None:
        d.mode = _cabi.Q_IDENTITY
        for name, t, dt in (("half_lo", half_lo, torch.float32), ("half_up", half_up, torch.float32), ("cell", cell, torch.int32)):
            if t is not None:
                _require_cuda(name, t, dt)
                if t.shape != z.shape:
                    raise ValueError(f"{name} must be shaped like z")
                tc = t.contiguous()
                keep.append(tc)
                setattr(d, name, tc.data_ptr())
                setattr(d, name + "_bs", Cc * hw)
        if g_dist is not None:
            if g_dist.dtype != torch.float64 or not g_dist.is_contiguous() or g_dist.device != z.device:
                raise ValueError("g_dist must be a contiguous float64 tensor on z's device")
            d.g_dist, d.n_dist = g_dist.data_ptr(), g_dist.numel()
    d.likelihood_bound = float(likelihood_bound)
    gm, gb, gf = [], [], []
    for i in range(5):
        m, b = matrices[i].detach().contiguous(), biases[i].detach().contiguous()
        keep += [m, b]
        d.matrix[i], d.bias[i] = m.data_ptr(), b.data_ptr()
        if need_params:
            gm.append(torch.empty_like(m)); gb.append(torch.empty_like(b))
            d.g_matrix[i], d.g_bias[i] = gm[-1].data_ptr(), gb[-1].data_ptr()
    for i in range(4):
        f = factors[i].detach().contiguous()
        keep.append(f)
        d.factor[i] = f.data_ptr()
        if need_params:
            gf.append(torch.empty_like(f))
            d.g_factor[i] = gf[-1].data_ptr()
    med = medians.detach().reshape(-1).contiguous()
    keep.append(med)
    d.medians = med.data_ptr()
    g_med = torch.zeros(Cc, dtype=torch.float32, device=z.device)
    d.g_medians = g_med.data_ptr()
    g_z = torch.empty_like(zc) if need_z else None
    if g_z is not None:
        d.g_z, d.g_z_bs = g_z.data_ptr(), Cc * hw
    d.philox_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    d.philox_offset = int(offset) & 0xFFFFFFFFFFFFFFFF
    ws = _eb_bwd_workspace(z.device, Cc)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(z.device):
        code = lib.reslic_eb_bwd_f32(C.byref(d), _cabi.current_stream_ptr(z.device))
    _cabi.check(code, "reslic_eb_bwd_f32")
    return g_z, gm, gb, gf, g_med


def lrp_tail_(y_hat: Tensor, lrp: Tensor) -> Tensor:
    """In place ``y_hat += 0.5 * tanh(lrp)`` (tcm.py:461-464) as one launch; y_hat may be a channel
    slice of the full latent."""
    lib = _cabi.load()
    _require_cuda("y_hat", y_hat)
    _require_cuda("lrp", lrp)
    if lrp.shape != y_hat.shape:
        raise ValueError("lrp shape must match y_hat")
    yh, ybs, n = image_major(y_hat)
    if yh is not y_hat:
        raise ValueError("y_hat must be image-major (a channel slice of a contiguous NCHW tensor qualifies)")
    lr, lbs, _ = image_major(lrp)
    B = y_hat.shape[0] if y_hat.dim() > 0 else 1
    with torch.cuda.device(y_hat.device):
        code = lib.reslic_lrp_tail_f32(yh.data_ptr(), ybs, lr.data_ptr(), lbs, B, n, _cabi.current_stream_ptr(y_hat.device))
    _cabi.check(code, "reslic_lrp_tail_f32")
    return y_hat
