"""The per-batch entropy pass as TCM drives it, minus the dense networks.

``TcmEntropyPath`` owns the two drop-in entropy models exactly as ``TCM.__init__`` does
(src/models/reference/tcm.py:416-417) and runs, for one batch of latents, what
``TCM.forward`` / ``TCM.compress`` run between the dense transforms (tcm.py:429-466,
527-552) and what ``RateDistortionLoss`` / ``compute_bpp`` reduce afterwards
(training/loss.py:24-27, eval.py:27-31):

    z  -> EntropyBottleneck:  z_hat (ste_round about the medians), L_z, bits_z[B]
    for each of the 5 channel slices of y (sequential in the real model: mu, sigma of
    slice k depend on y_hat of slices < k):
        GaussianConditional fused pass -> y_hat slice, L_y slice, [symbols, indexes], bits_y[B]
    bits[B] = bits_z + sum_k bits_y,k   (each launch adds its fixed-point sum to the rate workspace with
                                          fire-and-forget reductions; the last launch collects them into
                                          bits; bpp = bits / num_pixels)

= 1 + 5 kernel launches per batch and nothing else (no torch glue kernels).  Static output buffers make the pass CUDA-graph
capturable (``capture()``), which removes the Python/launch overhead from steady state.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _cabi, ops, synthetic
from .entropy_models import EntropyBottleneck, GaussianConditional


class TcmEntropyPath(nn.Module):
    # each slice launch prefetches the NEXT slice's y into L2 from the CTAs that finish early (reslic_gc_desc.next_y);
    # RESLIC_PREFETCH_NEXT=0 switches the hint off (A/B measurements)
    prefetch_next_slice = os.environ.get("RESLIC_PREFETCH_NEXT", "1") != "0"

    def __init__(self, z_channels: int = synthetic.Z_CHANNELS, num_slices: int = synthetic.NUM_SLICES):
        super().__init__()
        self.num_slices = int(num_slices)
        self.entropy_bottleneck = EntropyBottleneck(z_channels)      # tcm.py:416
        self.gaussian_conditional = GaussianConditional(None)        # tcm.py:417
        self._bufs: Optional[Dict[str, Tensor]] = None
        self._key = None

    # ------------------------------------------------------------------ static buffers
    def buffers(self, y: Tensor, z: Tensor, with_indexes: bool, training: bool) -> Dict[str, Tensor]:
        key = (tuple(y.shape), tuple(z.shape), bool(with_indexes), bool(training), y.device)
        if self._key != key:
            dev, B = y.device, y.shape[0]
            b = {
                "y_hat": torch.empty_like(y),
                "y_lik": torch.empty_like(y),
                "z_hat": torch.empty_like(z),
                "z_lik": torch.empty_like(z),
                "bits": torch.zeros(B, dtype=torch.float64, device=dev),
                "workspace": torch.zeros(max(int(_cabi.load().reslic_workspace_bytes(B)), 16),
                                         dtype=torch.uint8, device=dev),
            }
            if with_indexes:
                b["symbols"] = torch.empty(y.shape, dtype=torch.int32, device=dev)
                b["indexes"] = torch.empty(y.shape, dtype=torch.int32, device=dev)
            if training:
                b["y_noisy"] = torch.empty_like(y)
            self._bufs, self._key = b, key
        return self._bufs

    # ------------------------------------------------------------------ one pass
    @torch.no_grad()
    def forward(self, y: Tensor, mu: Tensor, sigma: Tensor, z: Tensor, *, training: bool = False,
                with_indexes: bool = False, num_pixels: Optional[int] = None, seed: int = 0,
                offset: int = 0, noise_y: Optional[Tensor] = None, noise_z: Optional[Tensor] = None,
                fuse_slices: bool = False, skip_z: bool = False, defer_rate: bool = False, exchange=None,
                exchange_step: int = 0, exchange_advance: bool = True) -> Dict[str, Tensor]:
        """All tensors on the GPU, NCHW fp32: y/mu/sigma [B, 320, h, w], z [B, 192, h/4, w/4].
        Returns views of static buffers (valid until the next call).

        ``fuse_slices`` runs all channels of y in ONE launch: what a model whose mu/sigma are
        available for every channel at once does (e.g. the reference's ScaleHyperprior,
        src/models/Balle2018.py: one gaussian_conditional call on the whole y); TCM itself needs
        the per-slice mode because slice k's parameters depend on y_hat of slices < k.
        ``defer_rate`` leaves the pass's rate in the workspace (no launch collects; ``bits`` is not
        written): the caller sums several passes and calls ``ops.rate_finalize`` once.
        ``exchange`` (a :class:`reslic_tcm_b200.dist.PeerRateExchange`): multi-GPU runs — the launch that collects the
        batch's rate also publishes it to every rank over NVLink; no collective kernel runs.  The batch
        is step ``exchange.cursor + exchange_step``; with ``exchange_advance`` the cursor moves on by one behind it (eager
        use) — a CUDA graph of n batches passes ``exchange_step=0..n-1, exchange_advance=False`` and calls
        ``exchange.advance(n)`` once."""
        gc, eb = self.gaussian_conditional, self.entropy_bottleneck
        b = self.buffers(y, z, with_indexes, training)
        C = y.shape[1]
        if C % self.num_slices:
            raise ValueError(f"{C} channels do not split into {self.num_slices} slices")
        n_launch = 1 if fuse_slices else self.num_slices
        cs = C // n_launch
        m, bi, f = eb._params()
        # Rate: every launch but the last adds its fixed-point sum to the workspace (fire-and-forget
        # reductions: the launch ends without the round trip that finding the last-arriving warp costs);
        # the LAST slice launch collects workspace + own sum into bits[b].  (The z launch is a poor
        # collector: 192 channel-warps per image arrive at once, 8.3 us vs 5.3 us deferred, while
        # collecting costs a slice launch 1.1 us.)
        if not skip_z:
            ops.eb_forward(z, m, bi, f, eb._medians_flat(), training=training, noise=noise_z,
                           likelihood_bound=eb._likelihood_bound if eb.use_likelihood_bound else 0.0,
                           want=("ste", "lik", "bits"),
                           out={"ste": b["z_hat"], "lik": b["z_lik"], "bits_deferred": True, "workspace": b["workspace"]},
                           seed=seed, offset=offset, lut=None if training else eb._eval_lut(),
                           next_y=y[:, :y.shape[1] // (1 if fuse_slices else self.num_slices)]
                           if (self.prefetch_next_slice and not fuse_slices) else None)
        want = ["ste", "lik", "bits"] + (["sym", "idx"] if with_indexes else []) + (["yhat"] if training else [])
        for k in range(n_launch):
            sl = slice(cs * k, cs * (k + 1))
            out = {"ste": b["y_hat"][:, sl], "lik": b["y_lik"][:, sl], "workspace": b["workspace"]}
            if k + 1 < n_launch or defer_rate:
                out["bits_deferred"] = True                                              # workspace += slice bits
            else:
                out["bits"], out["bits_collect"] = b["bits"], True                       # bits = slice + workspace
            if with_indexes:
                out["sym"], out["idx"] = b["symbols"][:, sl], b["indexes"][:, sl]
            if training:
                out["yhat"] = b["y_noisy"][:, sl]
            ops.gc_forward(y[:, sl], sigma[:, sl], mu[:, sl], training=training,
                           noise=None if noise_y is None else noise_y[:, sl],
                           scale_table=gc.scale_table if with_indexes else None, scale_bound=gc._scale_bound,
                           likelihood_bound=gc._likelihood_bound, want=want, out=out, seed=seed,
                           offset=offset + 1 + k,
                           next_y=y[:, cs * (k + 1):cs * (k + 2)] if (k + 1 < n_launch and self.prefetch_next_slice) else None,
                           exchange=exchange if (k + 1 == n_launch and not defer_rate) else None,
                           exchange_step=exchange_step)
        if exchange is not None and exchange_advance and not defer_rate:
            exchange.advance(1)
        res = {"y_hat": b["y_hat"], "z_hat": b["z_hat"], "bits": b["bits"],
               "likelihoods": {"y": b["y_lik"], "z": b["z_lik"]}}
        if with_indexes:
            res["symbols"], res["indexes"] = b["symbols"], b["indexes"]
        if training:
            res["y_noisy"] = b["y_noisy"]
        res["num_pixels"] = num_pixels      # bpp[b] = bits[b] / num_pixels (loss.py:24-27), left to the reader
        return res

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, y: Tensor, mu: Tensor, sigma: Tensor, z: Tensor, **kw):
        """Capture one pass over the given (static) input tensors into a CUDA graph.
        Returns (graph, result dict); refill the inputs in place and ``graph.replay()``.
        Training mode: kernel arguments are frozen at capture, so every replay of ONE graph draws the same Philox
        noise field (seed, offset are arguments).  A training loop that replays graphs captures a few of them with
        different ``offset=`` values and rotates (bench.py does: one per buffer set), or passes explicit
        ``noise_y`` / ``noise_z`` tensors that it refills; eager calls take ``offset`` per call."""
        self.buffers(y, z, kw.get("with_indexes", False), kw.get("training", False))
        side = torch.cuda.Stream(device=y.device)
        side.wait_stream(torch.cuda.current_stream(y.device))
        with torch.cuda.stream(side):      # warm-up outside capture (lazy module init, allocator)
            self.forward(y, mu, sigma, z, **kw)
        torch.cuda.current_stream(y.device).wait_stream(side)
        torch.cuda.synchronize(y.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            res = self.forward(y, mu, sigma, z, **kw)
        return graph, res


class TcmStanhEntropyPath(nn.Module):
    """The entropy pass of the reference's STanH model (src/models/stanh/tcm_stanh.py:331-336, 396-451): the plain
    factorized bottleneck on z (inherited from TCM, :403) and, per channel slice of y, one
    ``GaussianConditionalStanh`` call — annealed soft quantization about the predicted mean while ``training``
    (beta > 0), hard levels otherwise, with the variable-bin Gaussian likelihood (:432) — whose launch also writes
    ``ste_round(y - mu) + mu``, the value the slice loop carries on when the STanH is frozen (:433-434), and adds the
    slice's rate to the workspace; then ONE pass over the whole y for ``quantize(y, "training")`` +
    ``compute_gap`` (:448-449, 465-478: both squared-error sums in the same launch).

    = 1 + 5 + 1 kernel launches per batch over static buffers (CUDA-graph capturable, ``capture()``).  ``levels``
    mirrors the reference's one quantizer per lambda (``lv`` picks it)."""

    def __init__(self, gaussian_configuration: dict, z_channels: int = synthetic.Z_CHANNELS,
                 num_slices: int = synthetic.NUM_SLICES, levels: int = 1, channels: Optional[int] = None):
        super().__init__()
        from .stanh import GaussianConditionalStanh

        self.num_slices = int(num_slices)
        self.gaussian_configuration = dict(gaussian_configuration)
        self.entropy_bottleneck = EntropyBottleneck(z_channels)                                   # tcm.py:416
        self.gaussian_conditional = nn.ModuleList(                                                # tcm_stanh.py:331-336
            GaussianConditionalStanh(None, channels=z_channels if channels is None else channels,
                                     gaussian_configuration=self.gaussian_configuration) for _ in range(int(levels)))
        self._bufs: Optional[Dict[str, Tensor]] = None
        self._key = None

    def buffers(self, y: Tensor, z: Tensor) -> Dict[str, Tensor]:
        key = (tuple(y.shape), tuple(z.shape), y.device)
        if self._key != key:
            dev, B = y.device, y.shape[0]
            lib = _cabi.load()
            self._bufs = {
                "y_hat": torch.empty_like(y),        # ste_round(y - mu) + mu
                "y_q": torch.empty_like(y),          # the STanH output (soft while training, levels otherwise)
                "y_lik": torch.empty_like(y),
                "z_hat": torch.empty_like(z),
                "z_lik": torch.empty_like(z),
                "bits": torch.zeros(B, dtype=torch.float64, device=dev),
                "workspace": torch.zeros(max(int(lib.reslic_workspace_bytes(B)), 16), dtype=torch.uint8, device=dev),
                "gap_sums": torch.zeros(2, dtype=torch.float64, device=dev),
                "gap_workspace": torch.zeros(int(lib.reslic_stanh_gap_workspace_bytes()), dtype=torch.uint8, device=dev),
            }
            self._key = key
        return self._bufs

    @torch.no_grad()
    def forward(self, y: Tensor, mu: Tensor, sigma: Tensor, z: Tensor, *, training: bool = True, lv: int = 0,
                num_pixels: Optional[int] = None, seed: int = 0, offset: int = 0, noise_z: Optional[Tensor] = None,
                with_gap: bool = True, skip_z: bool = False, defer_rate: bool = False, **_ignored) -> Dict[str, Tensor]:
        """y/mu/sigma [B, C, h, w], z [B, Cz, h/4, w/4] on the GPU, fp32.  Returns views of static buffers:
        ``y_hat`` (ste values), ``y_q`` (STanH outputs), likelihoods, ``bits`` [B] and ``gap_sums``
        ([sum (y - soft)^2, sum (y - hard)^2]; gap = |difference| / y.numel(), tcm_stanh.py:465-478)."""
        if _ignored.get("exchange") is not None:
            raise ValueError("TcmStanhEntropyPath: the collecting launch is a STanH kernel, which does not publish; call "
                             "exchange.publish(res['bits']) behind the pass (bench.py: PeerExchange form 'branch')")
        if _ignored.get("with_indexes") or _ignored.get("fuse_slices"):
            raise ValueError("TcmStanhEntropyPath runs the training / evaluation forward per slice; symbols come from "
                             "GaussianConditionalStanh.quantize(..., 'symbols')")
        gc, eb = self.gaussian_conditional[lv], self.entropy_bottleneck
        gc.stanh.update_state(y.device)              # tcm_stanh.py:399 (a no-op while w, b are unchanged)
        b = self.buffers(y, z)
        C = y.shape[1]
        if C % self.num_slices:
            raise ValueError(f"{C} channels do not split into {self.num_slices} slices")
        cs = C // self.num_slices
        if not skip_z:
            m, bi, f = eb._params()
            ops.eb_forward(z, m, bi, f, eb._medians_flat(), training=training, noise=noise_z,
                           likelihood_bound=eb._likelihood_bound if eb.use_likelihood_bound else 0.0,
                           want=("ste", "lik", "bits"),
                           out={"ste": b["z_hat"], "lik": b["z_lik"], "bits_deferred": True, "workspace": b["workspace"]},
                           seed=seed, offset=offset, lut=None if training else eb._eval_lut())
        for k in range(self.num_slices):
            sl = slice(cs * k, cs * (k + 1))
            out = {"yhat": b["y_q"][:, sl], "lik": b["y_lik"][:, sl], "ste": b["y_hat"][:, sl], "workspace": b["workspace"]}
            if k + 1 < self.num_slices or defer_rate:
                out["bits_deferred"] = True
            else:
                out["bits"], out["bits_collect"] = b["bits"], True
            gc.forward_fused(y[:, sl], sigma[:, sl], training=training, means=mu[:, sl],
                             want=("yhat", "lik", "ste", "bits"), out=out)
        if with_gap:
            gc.stanh.gap_sums(y, out=b["gap_sums"], workspace=b["gap_workspace"])
        res = {"y_hat": b["y_hat"], "y_q": b["y_q"], "z_hat": b["z_hat"], "bits": b["bits"],
               "likelihoods": {"y": b["y_lik"], "z": b["z_lik"]}, "num_pixels": num_pixels}
        if with_gap:
            res["gap_sums"] = b["gap_sums"]
        return res

    def gap(self, res: Dict[str, Tensor], n: int) -> Tensor:
        """|MSE(y, soft) - MSE(y, hard)| from a forward's ``gap_sums`` (compute_gap, tcm_stanh.py:465-478)."""
        return torch.abs(res["gap_sums"][0] / n - res["gap_sums"][1] / n).to(torch.float32)

    def capture(self, y: Tensor, mu: Tensor, sigma: Tensor, z: Tensor, **kw):
        """As TcmEntropyPath.capture."""
        self.buffers(y, z)
        side = torch.cuda.Stream(device=y.device)
        side.wait_stream(torch.cuda.current_stream(y.device))
        with torch.cuda.stream(side):
            self.forward(y, mu, sigma, z, **kw)
        torch.cuda.current_stream(y.device).wait_stream(side)
        torch.cuda.synchronize(y.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            res = self.forward(y, mu, sigma, z, **kw)
        return graph, res


class HostPipeline:
    """End-to-end pass for latents that live in HOST memory: pinned host y/mu/sigma/z in, host
    symbols/indexes/bits out.  The batch is cut into image chunks (contiguous in NCHW); chunk c's
    H2D copy, its 1 + 5 kernel launches and its D2H copy run on three streams chained by events,
    so PCIe traffic in both directions overlaps and the kernels hide entirely under the copies.
    Device and host output buffers exist ``depth`` (2) times over, so consecutive ``run`` calls
    overlap too: batch i+1 is uploading while batch i's symbols are still going back, and the
    steady state is bounded by the larger of the two PCIe directions, not by their sum.
    This is what a caller whose dense transforms run elsewhere (or the CPU rANS coder consuming the
    symbols, tcm.py:551-565) sees; when the latents are produced on the GPU use TcmEntropyPath."""

    def __init__(self, path: TcmEntropyPath, batch: int, y_hw, z_hw, *, with_indexes: bool, training: bool = False,
                 chunks: int = 4, device=None, num_pixels: Optional[int] = None, depth: int = 2, seed: int = 0,
                 packed_slots: bool = False):
        self.path, self.with_indexes, self.training, self.num_pixels = path, with_indexes, training, num_pixels
        self.seed = int(seed)          # Philox key of the noise modes (the counter is the element index inside a launch)
        dev = torch.device(device if device is not None else "cuda")
        self.dev = dev
        B = batch
        chunks = max(1, min(chunks, B))
        base, extra = divmod(B, chunks)
        self.ranges = []
        start = 0
        for c in range(chunks):
            n = base + (1 if c < extra else 0)
            self.ranges.append((start, start + n))
            start += n
        self.depth = max(1, int(depth))
        C, Cz = synthetic.M_LATENT, synthetic.Z_CHANNELS
        f32 = dict(dtype=torch.float32, device=dev)
        # packed_slots: the rANS table lookup runs on the device (ops.rans_slots) and ONE packed (start, range) word per
        # symbol goes back instead of int32 symbols + int32 indexes — what rans.encode_slots_batch consumes
        self.packed_slots = bool(packed_slots)
        if self.packed_slots:
            gc = path.gaussian_conditional
            if not with_indexes or gc._quantized_cdf.numel() == 0:
                raise ValueError("packed_slots needs with_indexes and the CDF tables (update_scale_table) of the Gaussian conditional")
            self._tables = tuple(t.detach().to(device=dev, dtype=torch.int32).contiguous()
                                 for t in (gc._quantized_cdf, gc._cdf_length.reshape(-1), gc._offset.reshape(-1)))
        self.out_names = ["bits"] + (["symbols", "indexes"] if with_indexes else [])
        if self.packed_slots:
            self.out_names = ["bits", "slots", "esc_pos", "esc_raw", "slot_status"]
        self.slots = []
        for _ in range(self.depth):
            d_in = {k: torch.empty(B, C, *y_hw, **f32) for k in ("y", "mu", "sigma")}
            d_in["z"] = torch.empty(B, Cz, *z_hw, **f32)
            # one sub-path (static output buffers) per chunk, sharing the parameters of `path`
            sub = []
            for _r in self.ranges:
                p = TcmEntropyPath.__new__(TcmEntropyPath)
                nn.Module.__init__(p)
                p.num_slices = path.num_slices
                p.entropy_bottleneck, p.gaussian_conditional = path.entropy_bottleneck, path.gaussian_conditional
                p._bufs, p._key = None, None
                sub.append(p)
            h_out = {"bits": torch.empty(B, dtype=torch.float64).pin_memory()}
            d_slot = None
            if self.packed_slots:
                i32 = dict(dtype=torch.int32, device=dev)
                nchunk = len(self.ranges)
                caps = [(b - a) * C * y_hw[0] * y_hw[1] // 1024 + 1024 for a, b in self.ranges]   # escape-list capacity per chunk
                cap = max(caps)
                d_slot = [(torch.empty(b - a, C, *y_hw, **i32), torch.empty(cap, **i32),
                           torch.empty(cap, dtype=torch.int64, device=dev), torch.empty(2, **i32)) for a, b in self.ranges]
                h_out["slots"] = torch.empty(B, C, *y_hw, dtype=torch.int32).pin_memory()
                h_out["esc_pos"] = torch.empty(nchunk, cap, dtype=torch.int32).pin_memory()
                h_out["esc_raw"] = torch.empty(nchunk, cap, dtype=torch.int64).pin_memory()
                h_out["slot_status"] = torch.empty(nchunk, 2, dtype=torch.int32).pin_memory()
            elif with_indexes:
                for k in ("symbols", "indexes"):
                    h_out[k] = torch.empty(B, C, *y_hw, dtype=torch.int32).pin_memory()
            self.slots.append({
                "d_in": d_in, "sub": sub, "h_out": h_out, "used": False, "d_slot": d_slot,
                "ev_in": [torch.cuda.Event() for _r in self.ranges],      # chunk uploaded
                "ev_comp": [torch.cuda.Event() for _r in self.ranges],    # chunk computed (inputs free, outputs ready)
                "ev_out": [torch.cuda.Event() for _r in self.ranges],     # chunk downloaded (device outputs free)
                "done": torch.cuda.Event(),                               # whole batch on the host
            })
        self.s_h2d, self.s_comp, self.s_d2h = (torch.cuda.Stream(device=dev) for _ in range(3))
        self._turn = 0
        one = self.slots[0]
        self.h2d_bytes = sum(t.numel() * 4 for t in one["d_in"].values())
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in one["h_out"].values())

    @torch.no_grad()
    def run(self, host: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """host: pinned CPU tensors y, mu, sigma, z (full batch).  Enqueues the batch and returns its
        pinned host outputs plus ``"done"``, the event to synchronise before reading them; the buffers
        are reused ``depth`` calls later.  ``host`` must stay unchanged until the upload has run."""
        turn = self._turn
        slot = self.slots[turn % self.depth]
        self._turn += 1
        if not slot["used"]:      # first use: order after whatever the caller has enqueued so far
            cur = torch.cuda.current_stream(self.dev)
            for st in (self.s_h2d, self.s_comp, self.s_d2h):
                st.wait_stream(cur)
        d_in, h_out = slot["d_in"], slot["h_out"]
        for c, (a, b) in enumerate(self.ranges):
            with torch.cuda.stream(self.s_h2d):
                if slot["used"]:
                    self.s_h2d.wait_event(slot["ev_comp"][c])          # the previous tenant's kernels have read the inputs
                for k in ("y", "mu", "sigma", "z"):
                    d_in[k][a:b].copy_(host[k][a:b], non_blocking=True)
                slot["ev_in"][c].record(self.s_h2d)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(slot["ev_in"][c])
                if slot["used"]:
                    self.s_comp.wait_event(slot["ev_out"][c])          # ... and its outputs have left the device
                # Philox counter = (element index inside a launch, offset, seed): every launch of every chunk of every
                # run gets its own offset (forward() uses offset .. offset + 5), so no two draws share a noise field
                res = slot["sub"][c].forward(d_in["y"][a:b], d_in["mu"][a:b], d_in["sigma"][a:b], d_in["z"][a:b],
                                             training=self.training, with_indexes=self.with_indexes,
                                             num_pixels=self.num_pixels, seed=self.seed,
                                             offset=(turn * len(self.ranges) + c) * 8)
                if self.packed_slots:
                    ops.rans_slots(res["symbols"], res["indexes"], *self._tables, out=slot["d_slot"][c])
                slot["ev_comp"][c].record(self.s_comp)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(slot["ev_comp"][c])
                if self.packed_slots:
                    d_sl, d_ep, d_er, d_st = slot["d_slot"][c]
                    h_out["bits"][a:b].copy_(res["bits"], non_blocking=True)
                    h_out["slots"][a:b].copy_(d_sl, non_blocking=True)
                    h_out["esc_pos"][c].copy_(d_ep, non_blocking=True)
                    h_out["esc_raw"][c].copy_(d_er, non_blocking=True)
                    h_out["slot_status"][c].copy_(d_st, non_blocking=True)
                else:
                    for k in self.out_names:
                        h_out[k][a:b].copy_(res[k], non_blocking=True)
                slot["ev_out"][c].record(self.s_d2h)
        slot["done"].record(self.s_d2h)
        slot["used"] = True
        out = dict(h_out)
        out["done"] = slot["done"]
        return out

    def strings(self, out: Dict[str, Tensor], threads: Optional[int] = None):
        """rANS strings (one per image) of a ``packed_slots`` batch whose ``done`` event has been synchronised: the
        host coder's state update over the downloaded slots.  None if a chunk's escape list overflowed."""
        from . import rans

        res = []
        for c, (a, b) in enumerate(self.ranges):
            part = rans.encode_slots_batch(out["slots"][a:b], out["esc_pos"][c], out["esc_raw"][c], out["slot_status"][c],
                                           threads=threads)
            if part is None:
                return None
            res += part
        return res

    def synchronize(self) -> None:
        """Wait until every enqueued batch is on the host."""
        self.s_d2h.synchronize()
