"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the reference's OWN code.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden

The reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the reference itself:

* ``gc_golden.npz``  — TCM._likelihood / _standardized_cumulative, get_scale_table and
  ste_round EXECUTED FROM tcm.py's source (reference_shim.load_tcm_functions), plus the
  reference's unmodified ``GaussianConditionalStanh`` (default STanH parameters, for which
  its eval forward coincides with the CompressAI path away from ties/saturation) for
  forward() and build_indexes().
* ``eb_golden.npz``  — the reference's unmodified ``EntropyBottleneckStanh`` eval forward
  (default STanH parameters, medians 0): pins _logits_cumulative, the sign trick and the
  permute wrapper.

* ``stanh_update_golden.npz`` — the reference's unmodified ``GaussianConditionalStanh.update`` and
  ``EntropyBottleneckStanh.update`` (pmf / cdf / _offset / _cdf_length over the STanH levels): SURVEY.md §8f N2.

Third-party pieces that are NOT in the reference tree (compressai's quantize-about-medians,
GaussianConditional.update, pmf_to_quantized_cdf) have no runnable reference here; those
parts of the oracle are restated from the published algorithm and remain "unpinned".
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import reference_shim as shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _edge_vectors(table: torch.Tensor):
    """Edge cases of SURVEY.md §8c: ties, -0.0, sigma at/around every table value, below the
    bound, above the table, huge |y - mu| / sigma."""
    nxt = lambda t, d: torch.nextafter(t, torch.full_like(t, d))
    sig = torch.cat([table, nxt(table, math.inf), nxt(table, -math.inf),
                     torch.tensor([0.0, 0.01, 0.05, 0.10999, 0.11, 0.11001, 255.9, 256.0, 256.1, 300.0, 1e4])])
    d = torch.tensor([-0.0, 0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 0.49999997, 0.50000006, 3.0, -7.0, 40.0,
                      -40.0, 79.5, -80.5, 200.0, 1000.0, 1e-8, -1e-8])
    S, D = torch.meshgrid(sig, d, indexing="ij")
    mu = torch.zeros_like(S)
    mu[::2] = 0.375  # exactly representable offsets keep the ties exact
    mu[1::4] = -1.25
    y = D + mu
    return y.reshape(1, 1, -1, d.numel()).contiguous(), mu.reshape(1, 1, -1, d.numel()).contiguous(), \
        S.reshape(1, 1, -1, d.numel()).contiguous()


def gen_gc():
    t = shim.load_tcm_functions()
    em, _ = shim.load_stanh_modules()
    table = t.get_scale_table()
    g = torch.Generator().manual_seed(20261018)
    shape = (2, 64, 8, 8)
    mu = torch.randn(shape, generator=g)
    sigma = torch.exp(torch.empty(shape).uniform_(math.log(0.05), math.log(300.0), generator=g))
    y = mu + sigma * torch.randn(shape, generator=g)
    noise = torch.empty(shape).uniform_(-0.5, 0.5, generator=g)
    ey, emu, esig = _edge_vectors(table)

    cfg = dict(beta=10, num_sigmoids=0, extrema=80, trainable=True, removing_mean=True, symmetry=False)
    gcs = em.GaussianConditionalStanh(None, channels=64, gaussian_configuration=cfg)
    gcs.stanh.update_state(torch.device("cpu"))
    gcs.scale_table = table.clone()

    out = {"scale_table": table, "y": y, "mu": mu, "sigma": sigma, "noise": noise,
           "edge_y": ey, "edge_mu": emu, "edge_sigma": esig}
    with torch.no_grad():
        for tag, (yy, mm, ss) in {"": (y, mu, sigma), "edge_": (ey, emu, esig)}.items():
            ste = t.ste_round(yy - mm) + mm                                    # tcm.py:457
            out[tag + "ste"] = ste
            out[tag + "lik_unbounded"] = t.tcm._likelihood(ste, ss, mm)       # tcm.py:570-582
            out[tag + "indexes"] = gcs.build_indexes(ss)                       # adaptive_gaussian_conditional.py:606-617
        # noise-mode likelihood of y + u through the TCM twin
        out["lik_noise_unbounded"] = t.tcm._likelihood(y + noise, sigma, mu)
        # the reference's own STanH module, default parameters, eval mode; restricted to the
        # region where the hard STanH equals round(): |y - mu| < 79 and no ties
        yh, lik = gcs(y, sigma, training=False, means=mu)
        inside = ((y - mu).abs() < 79.0)
        out["stanh_yhat"], out["stanh_lik"], out["stanh_valid"] = yh, lik, inside
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "gc_golden.npz"),
                        **{k: v.numpy() for k, v in out.items()})
    return out


def gen_eb():
    em, _ = shim.load_stanh_modules()
    torch.manual_seed(77)
    cfg = dict(beta=10, num_sigmoids=0, extrema=30, trainable=True, symmetry=False)
    C = 8
    eb = em.EntropyBottleneckStanh(C, factorized_configuration=cfg)
    eb.stanh.update_state(torch.device("cpu"))
    g = torch.Generator().manual_seed(78)
    with torch.no_grad():
        for i in range(5):
            m = getattr(eb, f"_matrix{i}")
            m.add_(0.3 * torch.randn(m.shape, generator=g))
            if i < 4:
                f = getattr(eb, f"_factor{i}")
                f.copy_(0.5 * torch.randn(f.shape, generator=g))
    z = 3.0 * torch.randn((3, C, 4, 6), generator=g)
    # keep clear of ties (hard STanH yields half levels there) and inside +-extrema
    frac = z - torch.floor(z)
    z = torch.where((frac - 0.5).abs() < 1e-3, z + 0.01, z).clamp(-29.0, 29.0)
    with torch.no_grad():
        zhat, lik = eb(z, training=False)
    out = {"z": z, "zhat": zhat, "lik": lik}
    for i in range(5):
        out[f"_matrix{i}"] = getattr(eb, f"_matrix{i}").detach()
        out[f"_bias{i}"] = getattr(eb, f"_bias{i}").detach()
        if i < 4:
            out[f"_factor{i}"] = getattr(eb, f"_factor{i}").detach()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "eb_golden.npz"), **{k: v.numpy() for k, v in out.items()})
    return out


def gen_stanh():
    """The reference's own STanH modules (src/quantization/activation.py,
    src/entropy_models/adaptive_gaussian_conditional.py) on CPU: activation (hard / soft), forward
    in eval and training mode, quantize modes incl. the per-element "symbols" loop, and the
    compute_gap formula of src/models/stanh/tcm_stanh.py:465-478."""
    import contextlib
    import io

    import torch.nn.functional as F

    em, q = shim.load_stanh_modules()
    cases = {
        "A": dict(symmetry=False, extrema=80, beta=10, removing_mean=True, perturb=False),
        "B": dict(symmetry=False, extrema=10, beta=3, removing_mean=False, perturb=True),
        "C": dict(symmetry=True, extrema=6, beta=5, removing_mean=True, perturb=True),
    }
    out = {}
    for tag, c in cases.items():
        g = torch.Generator().manual_seed(900 + ord(tag))
        cfg = dict(beta=c["beta"], num_sigmoids=0, extrema=c["extrema"], trainable=True,
                   removing_mean=c["removing_mean"], symmetry=c["symmetry"])
        with contextlib.redirect_stdout(io.StringIO()):
            gcs = em.GaussianConditionalStanh(None, channels=8, gaussian_configuration=cfg)
            if c["perturb"]:
                with torch.no_grad():
                    gcs.stanh.w.mul_(1.0 + 0.2 * torch.rand(gcs.stanh.w.shape, generator=g))
            gcs.stanh.update_state(torch.device("cpu"))
            gcs.stanh.define_channels_map()
        shape = (2, 8, 6, 6)
        mu = torch.randn(shape, generator=g)
        sigma = torch.exp(torch.empty(shape).uniform_(math.log(0.05), math.log(40.0), generator=g))
        y = mu + torch.minimum(sigma, torch.tensor(0.4 * c["extrema"])) * torch.randn(shape, generator=g)
        # keep clear of exact ties with the thresholds (hard form is discontinuous there)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            st = gcs.stanh
            b_all = st.sym_b if c["symmetry"] else st.b
            w_all = st.sym_w if c["symmetry"] else st.w
            flat = y.reshape(1, 1, -1)
            hard = st(flat, -1).reshape(shape)
            soft = st(flat, c["beta"]).reshape(shape)
            gap = torch.abs(F.mse_loss(flat, st(flat, c["beta"])) - F.mse_loss(flat, st(flat, -1)))
            yh_eval, lik_eval = gcs(y, sigma, training=False, means=mu)
            yh_train, lik_train = gcs(y, sigma, training=True, means=mu)
            sym = gcs.quantize(y.clone(), "symbols", means=mu)
            lik_only = gcs._likelihood(yh_train, sigma, means=mu)
        for k, v in dict(y=y, mu=mu, sigma=sigma, w=w_all.detach(), b=torch.sort(b_all.detach())[0],
                         w_param=st.w.detach(), b_param=st.b.detach(),
                         cum_w=st.cum_w, avg=st.average_points, dist=st.distance_points, hard=hard, soft=soft,
                         gap=gap.reshape(1), yhat_eval=yh_eval, lik_eval=lik_eval, yhat_train=yh_train,
                         lik_train=lik_train, sym=sym, lik_unbounded_train=lik_only,
                         meta=torch.tensor([float(c["symmetry"]), c["extrema"], c["beta"], float(c["removing_mean"])])
                         ).items():
            out[f"{tag}_{k}"] = v.detach()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "stanh_golden.npz"), **{k: v.numpy() for k, v in out.items()})
    return out


def gen_eb_stanh():
    """The reference's own EntropyBottleneckStanh with perturbed STanH weights, eval and training."""
    import contextlib
    import io

    em, _ = shim.load_stanh_modules()
    out = {}
    for tag, sym in (("N", False), ("S", True)):
        torch.manual_seed(31 + int(sym))
        g = torch.Generator().manual_seed(32 + int(sym))
        cfg = dict(beta=4, num_sigmoids=0, extrema=6, trainable=True, symmetry=sym)
        C = 5
        with contextlib.redirect_stdout(io.StringIO()):
            eb = em.EntropyBottleneckStanh(C, factorized_configuration=cfg)
            with torch.no_grad():
                eb.stanh.w.mul_(1.0 + 0.2 * torch.rand(eb.stanh.w.shape, generator=g))
                for i in range(5):
                    m = getattr(eb, f"_matrix{i}")
                    m.add_(0.3 * torch.randn(m.shape, generator=g))
                    if i < 4:
                        f = getattr(eb, f"_factor{i}")
                        f.copy_(0.5 * torch.randn(f.shape, generator=g))
            eb.stanh.update_state(torch.device("cpu"))
        z = 2.5 * torch.randn((2, C, 3, 7), generator=g)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            zh_e, lik_e = eb(z, training=False)
            zh_t, lik_t = eb(z, training=True)
            st = eb.stanh
            b_all = st.sym_b if sym else st.b
            w_all = st.sym_w if sym else st.w
        d = dict(z=z, zhat_eval=zh_e, lik_eval=lik_e, zhat_train=zh_t, lik_train=lik_t, w=w_all.detach(),
                 b=torch.sort(b_all.detach())[0], w_param=st.w.detach(), b_param=st.b.detach(), cum_w=st.cum_w)
        for i in range(5):
            d[f"_matrix{i}"] = getattr(eb, f"_matrix{i}").detach()
            d[f"_bias{i}"] = getattr(eb, f"_bias{i}").detach()
            if i < 4:
                d[f"_factor{i}"] = getattr(eb, f"_factor{i}").detach()
        for k, v in d.items():
            out[f"{tag}_{k}"] = v.detach()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "eb_stanh_golden.npz"), **{k: v.numpy() for k, v in out.items()})
    return out


def gen_stanh_update():
    """N2 pin: the reference's OWN ``update()`` methods, unmodified, on CPU —
    ``GaussianConditionalStanh.update`` (src/entropy_models/adaptive_gaussian_conditional.py:397-454) and
    ``EntropyBottleneckStanh.update`` (src/entropy_models/adaptive_entropy_bottleneck.py:481-514): the per-scale /
    per-channel pmf over the STanH levels, the float cdf, ``_offset``, ``_cdf_length`` and ``_quantized_cdf``.
    ``_quantized_cdf`` passes through ``compressai._CXX.pmf_to_quantized_cdf``, a third-party C++ op that is NOT in the
    reference tree (the shim substitutes the oracle's restatement of the published algorithm), so that one array pins
    the reference's row assembly (``_pmf_to_cdf``, :197-205: pmf row + tail mass, padding) but not the op itself."""
    import contextlib
    import io

    em, _ = shim.load_stanh_modules()
    cpu = torch.device("cpu")
    out = {}
    cases = {
        "A": dict(symmetry=False, extrema=8, beta=10, perturb=False, table=[0.11, 0.5, 1.0, 4.0, 32.0]),
        "B": dict(symmetry=False, extrema=5, beta=3, perturb=True, table=[0.2, 0.9, 2.5, 11.0]),
        "C": dict(symmetry=True, extrema=6, beta=5, perturb=True, table=[0.11, 0.7, 3.0, 19.0, 64.0, 256.0]),
    }
    for tag, c in cases.items():
        g = torch.Generator().manual_seed(1700 + ord(tag))
        cfg = dict(beta=c["beta"], num_sigmoids=0, extrema=c["extrema"], trainable=True, removing_mean=True, symmetry=c["symmetry"])
        with contextlib.redirect_stdout(io.StringIO()):
            gcs = em.GaussianConditionalStanh(None, channels=4, gaussian_configuration=cfg)
            if c["perturb"]:
                with torch.no_grad():
                    gcs.stanh.w.mul_(1.0 + 0.2 * torch.rand(gcs.stanh.w.shape, generator=g))
            gcs.scale_table = torch.tensor(c["table"])
            gcs.update(cpu)
        st = gcs.stanh
        d = dict(scale_table=gcs.scale_table, w_param=st.w.detach(), b_param=st.b.detach(), cum_w=st.cum_w, pmf=gcs.pmf, cdf=gcs.cdf,
                 offset=torch.as_tensor(gcs._offset).reshape(-1), cdf_length=gcs._cdf_length.reshape(-1), quantized_cdf=gcs._quantized_cdf,
                 meta=torch.tensor([float(c["symmetry"]), c["extrema"], c["beta"]]))
        for k, v in d.items():
            out[f"gc{tag}_{k}"] = v.detach()
    for tag, sym in (("N", False), ("S", True)):
        torch.manual_seed(41 + int(sym))
        g = torch.Generator().manual_seed(42 + int(sym))
        cfg = dict(beta=4, num_sigmoids=0, extrema=6, trainable=True, symmetry=sym)
        C = 5
        with contextlib.redirect_stdout(io.StringIO()):
            eb = em.EntropyBottleneckStanh(C, factorized_configuration=cfg)
            with torch.no_grad():
                eb.stanh.w.mul_(1.0 + 0.2 * torch.rand(eb.stanh.w.shape, generator=g))
                for i in range(5):
                    m = getattr(eb, f"_matrix{i}")
                    m.add_(0.3 * torch.randn(m.shape, generator=g))
                    if i < 4:
                        f = getattr(eb, f"_factor{i}")
                        f.copy_(0.5 * torch.randn(f.shape, generator=g))
            eb.update(cpu)
        d = dict(w_param=eb.stanh.w.detach(), b_param=eb.stanh.b.detach(), cum_w=eb.stanh.cum_w, pmf=eb.pmf.detach(), cdf=eb.cdf.detach())
        for i in range(5):
            d[f"_matrix{i}"] = getattr(eb, f"_matrix{i}").detach()
            d[f"_bias{i}"] = getattr(eb, f"_bias{i}").detach()
            if i < 4:
                d[f"_factor{i}"] = getattr(eb, f"_factor{i}").detach()
        for k, v in d.items():
            out[f"eb{tag}_{k}"] = v.detach()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "stanh_update_golden.npz"), **{k: v.numpy() for k, v in out.items()})
    return out


if __name__ == "__main__":
    if not shim.available():
        raise SystemExit("reference not available: golden vectors can only be generated in the build container")
    a = gen_gc()
    b = gen_eb()
    c = gen_stanh()
    gen_eb_stanh()
    u = gen_stanh_update()
    print("stanh:", len(c), "arrays; stanh update:", len(u), "arrays")
    print("wrote", GOLDEN_DIR, {k: tuple(v.shape) for k, v in a.items()}, {k: tuple(v.shape) for k, v in b.items()})
