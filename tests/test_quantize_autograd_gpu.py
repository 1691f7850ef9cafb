"""GPU: the quantize-only entry points under autograd.  The unmodified reference calls them with grad
enabled — ``self.gaussian_conditional[lv].quantize(y, mode="training")`` in src/models/stanh/tcm_stanh.py:448 and
wacnn_stanh.py:319, ``quantize(y, "noise")`` inside GainBalle2018's training forward — so a drop-in must not raise
there and must return the gradient autograd gives the reference's op sequence."""
import pytest
import torch

from oracle import stanh_ref as sr
from reslic_tcm_b200 import EntropyBottleneck, GaussianConditional, _cabi, stanh

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_compressai_quantize_modes_under_grad():
    g = torch.Generator().manual_seed(3)
    y = (3 * torch.randn(2, 8, 4, 4, generator=g)).to(DEV).requires_grad_(True)
    mu = torch.randn(2, 8, 4, 4, generator=g).to(DEV).requires_grad_(True)
    wy = torch.randn(2, 8, 4, 4, generator=g).to(DEV)
    gc = GaussianConditional(None).to(DEV)
    out = gc.quantize(y, "noise", mu)                       # y + U(-1/2, 1/2); means ignored (App. A.1)
    assert bool(((out - y).abs() <= 0.5).all())
    (out * wy).sum().backward()
    assert torch.equal(y.grad, wy) and mu.grad is None
    y.grad = None
    out = gc.quantize(y, "dequantize", mu)                  # round(y - mu) + mu
    assert torch.equal(out.detach(), torch.round(y.detach() - mu.detach()) + mu.detach())
    (out * wy).sum().backward()
    assert torch.equal(y.grad, torch.zeros_like(y)) and torch.equal(mu.grad, wy)     # what torch.round's zero gradient leaves
    sym = gc.quantize(y, "symbols", mu)
    assert sym.dtype == torch.int32 and not sym.requires_grad
    # the bottleneck's medians are a Parameter: quantize about them inside a grad-enabled region
    eb = EntropyBottleneck(8).to(DEV)
    z = torch.randn(2, 8, 3, 3, device=DEV)
    med = eb._get_medians().detach().reshape(1, -1, 1, 1).expand_as(z).contiguous().requires_grad_(True)
    zq = eb.quantize(z, "dequantize", med)
    zq.sum().backward()
    assert torch.equal(med.grad, torch.ones_like(med))


@pytest.mark.parametrize("mode,removing_mean,symmetry", [("training", True, False), ("training", False, False),
                                                         ("training", True, True), ("dequantize", True, False)])
def test_stanh_quantize_under_grad_matches_autograd(mode, removing_mean, symmetry):
    beta, extrema = 3.0, 6
    cfg = dict(beta=beta, num_sigmoids=0, extrema=extrema, trainable=True, removing_mean=removing_mean, symmetry=symmetry)
    mod = stanh.GaussianConditionalStanh(None, channels=4, gaussian_configuration=cfg).to(DEV)
    g = torch.Generator().manual_seed(43)
    with torch.no_grad():
        mod.stanh.w.mul_((1.0 + 0.2 * torch.rand(mod.stanh.w.shape, generator=g)).to(DEV))
    mod.stanh.update_state(torch.device(DEV))
    st = mod.stanh
    w = (st.sym_w if symmetry else st.w).detach().cpu()
    b = torch.sort((st.sym_b if symmetry else st.b).detach().cpu())[0]
    shape = (2, 4, 5, 6)
    mu = torch.randn(shape, generator=g)
    y = mu + 2.0 * torch.randn(shape, generator=g)
    wy = torch.randn(shape, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (y, mu)]
    ref = sr.quantize(leaves[0], mode, leaves[1], w, b, beta, symmetry, removing_mean)
    (ref * wy).sum().backward()

    dl = [t.clone().to(DEV).requires_grad_(True) for t in (y, mu)]
    assert torch.is_grad_enabled() and st.w.requires_grad          # exactly the situation that used to raise
    out = mod.quantize(dl[0], mode, means=dl[1])
    assert torch.allclose(out.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-5)
    (out * wy.to(DEV)).sum().backward()
    for name, a, r in zip(("d/dy", "d/dmu"), (t.grad for t in dl), (t.grad for t in leaves)):
        r = torch.zeros(shape) if r is None else r
        a = torch.zeros(shape) if a is None else a.cpu()
        scale = max(r.abs().max().item(), 1.0)
        assert bool(((a - r).abs() <= 5e-4 * r.abs() + 2e-5 * scale).all()), f"{mode} {name}: max err {(a - r).abs().max():.3g}"
    if mode == "training":
        assert st.w.grad is not None and bool(torch.isfinite(st.w.grad).all()) and float(st.w.grad.abs().sum()) > 0
    sym = mod.quantize(dl[0], "symbols", means=dl[1])
    assert sym.dtype == torch.int32 and not sym.requires_grad


def test_stanh_activation_and_compute_gap_path_under_grad():
    """tcm_stanh.py:448 + :465-478: y_gap = quantize(y, "training") then the gap, with grad enabled and trainable w, b."""
    cfg = dict(beta=5.0, num_sigmoids=0, extrema=8, trainable=True, removing_mean=True, symmetry=False)
    mod = stanh.GaussianConditionalStanh(None, channels=4, gaussian_configuration=cfg).to(DEV)
    mod.stanh.update_state(torch.device(DEV))
    y = (3 * torch.randn(2, 4, 6, 6, device=DEV)).requires_grad_(True)
    y_gap = mod.quantize(y, mode="training")                        # no means, grad enabled: must not raise
    soft = mod.stanh(y.reshape(1, 1, -1), mod.stanh.beta)           # the bare activation, same conditions
    assert torch.allclose(soft.reshape_as(y_gap), y_gap, atol=1e-6)
    gap = stanh.compute_gap(mod.stanh, y.detach())
    assert gap.ndim == 0 and float(gap) >= 0
    y_gap.sum().backward()
    assert y.grad is not None and bool((y.grad >= 0).all())          # a sum of increasing tanh steps


def test_entropy_bottleneck_stanh_forward_trains_and_the_fused_extension_refuses_grad():
    """forward() records an autograd node (tests/test_eb_stanh_backward_gpu.py checks the gradients); the multi-output
    forward_fused() extension has no backward and refuses to cut the graph silently."""
    cfg = dict(beta=4, num_sigmoids=0, extrema=6, trainable=True, symmetry=False)
    eb = stanh.EntropyBottleneckStanh(5, factorized_configuration=cfg).to(DEV)
    eb.stanh.update_state(torch.device(DEV))
    z = torch.randn(2, 5, 3, 3, device=DEV)
    zh, lik = eb(z, training=True)                                  # the reference's call, grad enabled, trainable parameters
    (zh.sum() + torch.log(lik).sum()).backward()
    assert eb._matrix0.grad is not None and eb.stanh.w.grad is not None and bool(torch.isfinite(eb.stanh.w.grad).all())
    with pytest.raises(_cabi.ReslicError, match="no autograd graph"):
        eb.forward_fused(z, training=True)
    with torch.no_grad():
        r = eb.forward_fused(z, training=True)
    assert torch.equal(r["zhat"], zh.detach()) and torch.equal(r["lik"], lik.detach())
    zq = eb.quantize(z.requires_grad_(True), "training")            # the quantizer alone is differentiable too
    zq.sum().backward()
    assert z.grad is not None and bool(torch.isfinite(z.grad).all())
