"""GPU: EntropyBottleneckStanh.forward under autograd (the class is called inside the training forward of
src/models/stanh/wacnn_stanh.py:160 and balle18_stanh.py:26,124) — gradients w.r.t. z, the bottleneck's parameters and the
trainable STanH weights against torch autograd through the oracle's op sequence (oracle/stanh_ref.eb_stanh_forward,
itself pinned to the reference's module by tests/test_stanh_oracle_golden.py)."""
import pytest
import torch

from oracle import compressai_ref as cr
from oracle import stanh_ref as sr
from reslic_tcm_b200 import stanh

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, r, what, rtol=2e-3, atol_scale=5e-5):
    a, r = a.double().cpu(), r.double()
    scale = max(float(r.abs().max()), 1e-6)
    err = (a - r).abs()
    assert bool((err <= rtol * r.abs() + atol_scale * scale).all()), f"{what}: max err {float(err.max()):.3g} (scale {scale:.3g})"


@pytest.mark.parametrize("training,symmetry", [(True, False), (True, True), (False, False)])
def test_eb_stanh_forward_backward_matches_autograd(training, symmetry):
    torch.manual_seed(5)
    g = torch.Generator().manual_seed(77 + int(symmetry))
    C, beta = 5, 4.0
    cfg = dict(beta=beta, num_sigmoids=0, extrema=6, trainable=True, symmetry=symmetry)
    mod = stanh.EntropyBottleneckStanh(C, factorized_configuration=cfg).to(DEV)
    with torch.no_grad():
        mod.stanh.w.mul_((1.0 + 0.2 * torch.rand(mod.stanh.w.shape, generator=g)).to(DEV))
        for i in range(5):
            m_ = getattr(mod, f"_matrix{i}")
            m_.add_((0.3 * torch.randn(m_.shape, generator=g)).to(DEV))
            if i < 4:
                f_ = getattr(mod, f"_factor{i}")
                f_.copy_((0.5 * torch.randn(f_.shape, generator=g)).to(DEV))
    mod.stanh.update_state(torch.device(DEV))
    z = 2.5 * torch.randn((3, C, 4, 7), generator=g)
    wz, wl = torch.randn(z.shape, generator=g), torch.randn(z.shape, generator=g)

    # ---- oracle with torch autograd (CPU, fp64 for a clean reference)
    ref = cr.EntropyBottleneckRef(C)
    leaves = {}
    for i in range(5):
        leaves[f"m{i}"] = getattr(mod, f"_matrix{i}").detach().cpu().double().requires_grad_(True)
        leaves[f"b{i}"] = getattr(mod, f"_bias{i}").detach().cpu().double().requires_grad_(True)
        if i < 4:
            leaves[f"f{i}"] = getattr(mod, f"_factor{i}").detach().cpu().double().requires_grad_(True)
    ref.matrices = [leaves[f"m{i}"] for i in range(5)]
    ref.biases = [leaves[f"b{i}"] for i in range(5)]
    ref.factors = [leaves[f"f{i}"] for i in range(4)]
    ref.quantiles = torch.zeros(C, 1, 3, dtype=torch.float64)
    w_raw = mod.stanh.w.detach().cpu().double().requires_grad_(True)
    b_raw = mod.stanh.b.detach().cpu().double().requires_grad_(True)
    w_eff, b_sorted, _ = mod.stanh._effective(w_raw, b_raw)          # the module's own (differentiable) map raw -> effective
    half = torch.cat((w_raw.new_zeros(1), torch.cumsum(w_raw, dim=0)))     # the levels as differentiable functions of w
    cum_w = torch.cat((-torch.flip(half[1:], dims=[0]), half), dim=0) if symmetry else half - w_raw.sum() / 2    # activation.py:91-98, 214-234
    assert torch.allclose(cum_w.detach().float(), mod.stanh.cum_w.cpu(), atol=1e-5)
    zl = z.double().requires_grad_(True)
    zh_r, lik_r = sr.eb_stanh_forward(zl, ref, w_eff, b_sorted, cum_w, beta, symmetry, training)
    ((zh_r * wz.double()).sum() + (torch.log(lik_r) * wl.double()).sum()).backward()

    # ---- the drop-in module under autograd
    zd = z.to(DEV).requires_grad_(True)
    zh, lik = mod(zd, training=training)
    assert torch.allclose(zh.detach().cpu().double(), zh_r.detach(), atol=2e-5) and torch.allclose(lik.detach().cpu().double(), lik_r.detach(), rtol=2e-4, atol=1e-7)
    ((zh * wz.to(DEV)).sum() + (torch.log(lik) * wl.to(DEV)).sum()).backward()
    if training:
        _close(zd.grad, zl.grad, "d/dz")
    else:
        assert zd.grad is None or float(zd.grad.abs().max()) == 0.0          # hard quantizer: zero gradient, as sign / relu give
    for i in range(5):
        _close(getattr(mod, f"_matrix{i}").grad, leaves[f"m{i}"].grad, f"d/d_matrix{i}")
        _close(getattr(mod, f"_bias{i}").grad, leaves[f"b{i}"].grad, f"d/d_bias{i}")
        if i < 4:
            _close(getattr(mod, f"_factor{i}").grad, leaves[f"f{i}"].grad, f"d/d_factor{i}")
    if training:
        _close(mod.stanh.w.grad, w_raw.grad, "d/d stanh.w", rtol=5e-3, atol_scale=2e-4)
        if b_raw.grad is not None:
            _close(mod.stanh.b.grad, b_raw.grad, "d/d stanh.b", rtol=5e-3, atol_scale=2e-4)
