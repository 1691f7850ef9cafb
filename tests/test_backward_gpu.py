"""GPU: fused backward kernels against torch autograd through the oracle's op sequence (CPU)."""
import pytest
import torch

from oracle import compressai_ref as cr
from oracle.reference_shim import LowerBound as RefLowerBound   # autograd rule of compressai.ops.LowerBound
from reslic_tcm_b200 import GaussianConditional

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref_forward(y, sigma, mu, noise, training):
    """compressai GaussianConditional.forward + ste_round with autograd (reference op order)."""
    lb_s, lb_l = RefLowerBound(0.11), RefLowerBound(1e-9)
    if training:
        outputs = y + noise
    else:
        outputs = torch.round(y - mu) + mu
    values = torch.abs(outputs - mu)
    s = lb_s(sigma)
    upper = cr.standardized_cumulative((0.5 - values) / s)
    lower = cr.standardized_cumulative((-0.5 - values) / s)
    lik = lb_l(upper - lower)
    ste = cr.ste_round(y - mu) + mu
    return outputs, lik, ste


@pytest.mark.parametrize("training", [True, False])
def test_gc_backward_matches_autograd(training):
    g = torch.Generator().manual_seed(17 + int(training))
    shape = (2, 16, 8, 8)
    mu = torch.randn(shape, generator=g)
    sigma = torch.exp(torch.empty(shape).uniform_(-3.0, 3.0, generator=g))   # some below the 0.11 bound
    y = mu + sigma * torch.randn(shape, generator=g) * 1.5
    y.view(-1)[:4] = mu.view(-1)[:4] + torch.tensor([40.0, -40.0, 0.0, 0.5])  # floor-bounded likelihoods, |.|'(0)
    noise = torch.empty(shape).uniform_(-0.5, 0.5, generator=g)
    wy, wl, ws = (torch.randn(shape, generator=g) for _ in range(3))

    leaves = [t.clone().requires_grad_(True) for t in (y, sigma, mu)]
    out, lik, ste = _ref_forward(*leaves[:1], leaves[1], leaves[2], noise, training) if False else \
        _ref_forward(leaves[0], leaves[1], leaves[2], noise, training)
    loss = (out * wy).sum() + (torch.log(lik) * wl).sum() + (ste * ws).sum()
    loss.backward()
    ref = [t.grad for t in leaves]

    gc = GaussianConditional(None).to(DEV)
    dl = [t.clone().to(DEV).requires_grad_(True) for t in (y, sigma, mu)]
    o2, l2, s2 = gc.forward_with_ste(dl[0], dl[1], dl[2], training=training, noise=noise.to(DEV) if training else None)
    loss2 = (o2 * wy.to(DEV)).sum() + (torch.log(l2) * wl.to(DEV)).sum() + (s2 * ws.to(DEV)).sum()
    loss2.backward()
    for name, a, r in zip(("d/dy", "d/dsigma", "d/dmu"), (t.grad for t in dl), ref):
        a = a.cpu()
        scale = r.abs().max().item()
        err = (a - r).abs()
        tol = 2e-5 * r.abs() + 2e-6 * max(scale, 1.0)
        assert bool((err <= tol).all()), f"{name}: max err {err.max():.3g} (scale {scale:.3g})"


def test_module_forward_is_differentiable_and_philox_backward_is_consistent():
    torch.manual_seed(0)
    gc = GaussianConditional(None).to(DEV).train()
    y = torch.randn(2, 8, 4, 4, device=DEV, requires_grad=True)
    s = (torch.rand(2, 8, 4, 4, device=DEV) + 0.3).requires_grad_(True)
    m = torch.randn(2, 8, 4, 4, device=DEV, requires_grad=True)
    y_hat, lik = gc(y, s, m)                     # training mode, in-kernel Philox noise
    u = (y_hat - y).detach()
    (-(torch.log2(lik)).sum() + (y_hat ** 2).sum()).backward()
    # same noise re-fed explicitly must give the same gradients (the backward regenerates it)
    y2, s2, m2 = (t.detach().clone().requires_grad_(True) for t in (y, s, m))
    yh2, lik2, _ = gc.forward_with_ste(y2, s2, m2, training=True, noise=u)
    (-(torch.log2(lik2)).sum() + (yh2 ** 2).sum()).backward()
    for a, b in ((y.grad, y2.grad), (s.grad, s2.grad), (m.grad, m2.grad)):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
