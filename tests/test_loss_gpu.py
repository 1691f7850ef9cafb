"""GPU: the RateDistortionLoss mirror (src/training/loss.py:6-35) and the likelihood-tensor rate reduction against
the reference's op sequence (torch.log(L).sum() / (-ln 2 * num_pixels) + MSE) with autograd."""
import math

import pytest
import torch

from reslic_tcm_b200 import ops
from reslic_tcm_b200.loss import RateDistortionLoss, rate_bits

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference_loss(output, target, lmbda):
    N, _, H, W = target.size()
    num_pixels = N * H * W
    bpp = sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in output["likelihoods"].values())
    mse = torch.nn.functional.mse_loss(output["x_hat"], target)
    return {"bpp_loss": bpp, "mse_loss": mse, "loss": lmbda * 255 ** 2 * mse + bpp}


@pytest.mark.parametrize("shape_y", [(3, 320, 16, 8), (2, 5, 7, 3)])        # 128-bit path / scalar tail and odd sizes
def test_loss_matches_reference_formula_and_gradients(shape_y):
    g = torch.Generator().manual_seed(8)
    B = shape_y[0]
    ly = (torch.rand(shape_y, generator=g) * 0.999 + 1e-9).double()
    lz = (torch.rand(B, 192, 4, 2, generator=g) * 0.5 + 1e-6).double()
    x = torch.rand(B, 3, 64, 32, generator=g).double()
    xh = (x + 0.05 * torch.randn(x.shape, generator=g).double())
    ref_in = [t.clone().requires_grad_(True) for t in (ly, lz, xh)]
    ref = _reference_loss({"likelihoods": {"y": ref_in[0], "z": ref_in[1]}, "x_hat": ref_in[2]}, x, 0.013)
    ref["loss"].backward()
    ours_in = [t.float().to(DEV).requires_grad_(True) for t in (ly, lz, xh)]
    crit = RateDistortionLoss(lmbda=[0.013]).to(DEV)
    out = crit({"likelihoods": {"y": ours_in[0], "z": ours_in[1]}, "x_hat": ours_in[2]}, x.float().to(DEV))
    out["loss"].backward()
    for k in ("bpp_loss", "mse_loss", "loss"):
        assert out[k].dtype == torch.float32
        assert float(out[k].detach()) == pytest.approx(float(ref[k].detach()), rel=1e-5), k
    for a, r, name in zip(ours_in, ref_in, ("d/dLy", "d/dLz", "d/dx_hat")):
        # (fp32 inputs against the fp64 formula: x_hat - target cancels, so the tolerance has an absolute part)
        err = (a.grad.cpu().double() - r.grad).abs()
        assert bool((err <= 2e-5 * r.grad.abs() + 1e-6 * r.grad.abs().max()).all()), (name, float(err.max()))
    with pytest.raises(NotImplementedError):
        RateDistortionLoss(type="ms_ssim")({"likelihoods": {"y": ours_in[0]}, "x_hat": ours_in[2]}, x.float().to(DEV))


def test_rate_from_likelihood_modes_slices_and_reproducibility():
    torch.manual_seed(3)
    lik = (torch.rand(4, 320, 12, 8, device=DEV) * 0.9 + 1e-7)
    want = -(torch.log2(lik.double()).reshape(4, -1).sum(1))
    a = ops.rate_from_likelihood(lik)
    assert torch.allclose(a, want, rtol=2e-6) and torch.equal(a, ops.rate_from_likelihood(lik))     # bit-reproducible
    sl = lik[:, 64:128]                                      # a channel slice: image-major view, no copy
    assert torch.allclose(ops.rate_from_likelihood(sl), -(torch.log2(sl.double()).reshape(4, -1).sum(1)), rtol=2e-6)
    acc = a.clone()
    ops.rate_from_likelihood(lik, out={"bits": acc, "bits_accumulate": True})
    assert torch.allclose(acc, 2 * want, rtol=2e-6)
    bad = lik.clone()
    bad[1, 0, 0, 0] = float("nan")
    bad[2, 0, 0, 0] = 0.0
    r = ops.rate_from_likelihood(bad)
    assert torch.isnan(r[1]) and torch.isinf(r[2]) and torch.allclose(r[[0, 3]], want[[0, 3]], rtol=2e-6)
    assert torch.allclose(rate_bits(lik), want, rtol=2e-6)
    assert ops.rate_from_likelihood(torch.empty(0, 4, device=DEV)).numel() == 0
