"""GPU: the drop-in modules inside a TCM-shaped slice loop WITH dense layers and a training step —
forward values and parameter gradients against the same loop built from the oracle's ops under
torch autograd on the CPU.  Mirrors the structure of TCM.forward (src/models/reference/tcm.py:425-478)
at toy size: hyper-prior -> per-slice (mu, sigma) nets fed by previous y_hat slices -> entropy models
-> rate-distortion loss (src/training/loss.py:16-35)."""
import math

import pytest
import torch
import torch.nn as nn

from oracle import compressai_ref as cr
from oracle.reference_shim import LowerBound as RefLowerBound
from reslic_tcm_b200 import EntropyBottleneck, GaussianConditional, ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
S, CS = 3, 4          # slices, channels per slice


class MiniTCM(nn.Module):
    def __init__(self, use_cuda_path: bool):
        super().__init__()
        torch.manual_seed(7)
        self.use_cuda_path = use_cuda_path
        self.h_a = nn.Conv2d(S * CS, 6, 3, stride=2, padding=1)
        self.h_mean = nn.ConvTranspose2d(6, S * CS, 4, stride=2, padding=1)
        self.h_scale = nn.ConvTranspose2d(6, S * CS, 4, stride=2, padding=1)
        self.cc_mean = nn.ModuleList(nn.Conv2d(S * CS + CS * k, CS, 3, padding=1) for k in range(S))
        self.cc_scale = nn.ModuleList(nn.Conv2d(S * CS + CS * k, CS, 3, padding=1) for k in range(S))
        self.lrp = nn.ModuleList(nn.Conv2d(S * CS + CS * (k + 1), CS, 3, padding=1) for k in range(S))
        if use_cuda_path:
            self.entropy_bottleneck = EntropyBottleneck(6)
            self.gaussian_conditional = GaussianConditional(None)
        else:
            self.ref_eb = cr.EntropyBottleneckRef(6)

    def _eb(self, z, noise):
        if self.use_cuda_path:
            return self.entropy_bottleneck(z, training=True, noise=noise)
        eb = self.ref_eb
        xp = z.permute(1, 0, 2, 3).reshape(6, 1, -1)
        out = xp + noise.permute(1, 0, 2, 3).reshape(6, 1, -1)
        lik = RefLowerBound(1e-9)(eb._likelihood(out))
        back = lambda t: t.reshape(6, z.shape[0], z.shape[2], z.shape[3]).permute(1, 0, 2, 3)
        return back(out), back(lik)

    def _gc(self, y, s, m, noise):
        if self.use_cuda_path:
            return self.gaussian_conditional.forward_with_ste(y, s, m, training=True, noise=noise)
        out = y + noise
        values = torch.abs(out - m)
        sb = RefLowerBound(0.11)(s)
        lik = RefLowerBound(1e-9)(cr.standardized_cumulative((0.5 - values) / sb) -
                                  cr.standardized_cumulative((-0.5 - values) / sb))
        return out, lik, cr.ste_round(y - m) + m

    def forward(self, y, noise_y, noise_z):
        z = self.h_a(y)
        _, z_lik = self._eb(z, noise_z)
        med = (self.entropy_bottleneck._get_medians() if self.use_cuda_path else self.ref_eb._get_medians())
        med = med.reshape(1, -1, 1, 1)
        z_hat = cr.ste_round(z - med) + med
        means, scales = self.h_mean(z_hat), torch.exp(0.5 * self.h_scale(z_hat))
        y_hat_slices, liks = [], []
        for k, y_s in enumerate(y.chunk(S, 1)):
            sup = torch.cat([means] + y_hat_slices, 1)
            mu = self.cc_mean[k](sup)
            sigma = torch.exp(0.3 * self.cc_scale[k](torch.cat([scales] + y_hat_slices, 1)))
            _, lik, y_hat = self._gc(y_s, sigma, mu, noise_y[:, k * CS:(k + 1) * CS])
            liks.append(lik)
            y_hat = y_hat + 0.5 * torch.tanh(self.lrp[k](torch.cat([sup, y_hat], 1)))      # tcm.py:461-464
            y_hat_slices.append(y_hat)
        y_hat = torch.cat(y_hat_slices, 1)
        num_pixels = y.shape[0] * (16 * y.shape[2]) * (16 * y.shape[3])
        bpp = sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in (torch.cat(liks, 1), z_lik))
        mse = ((y_hat - y) ** 2).mean()
        return bpp + 0.01 * 255 ** 2 * mse, bpp


def test_training_step_matches_cpu_reference_loop():
    g = torch.Generator().manual_seed(11)
    y = 3.0 * torch.randn(2, S * CS, 8, 8, generator=g)
    noise_y = torch.empty(y.shape).uniform_(-0.5, 0.5, generator=g)
    noise_z = torch.empty(2, 6, 4, 4).uniform_(-0.5, 0.5, generator=g)
    ref = MiniTCM(False)
    ours = MiniTCM(True)
    # same dense weights and same bottleneck parameters on both sides
    ours.load_state_dict({k: v for k, v in ref.state_dict().items()}, strict=False)
    eb_params = synthetic.eb_parameters(6, trained_like=True, seed=3)
    synthetic.load_eb_parameters(ours.entropy_bottleneck, eb_params)
    ref.ref_eb.matrices = [eb_params[f"_matrix{i}"].clone().requires_grad_(True) for i in range(5)]
    ref.ref_eb.biases = [eb_params[f"_bias{i}"].clone().requires_grad_(True) for i in range(5)]
    ref.ref_eb.factors = [eb_params[f"_factor{i}"].clone().requires_grad_(True) for i in range(4)]
    ref.ref_eb.quantiles = eb_params["quantiles"].clone()
    ours = ours.to(DEV).train()

    loss_r, bpp_r = ref(y, noise_y, noise_z)
    loss_r.backward()
    loss_o, bpp_o = ours(y.to(DEV), noise_y.to(DEV), noise_z.to(DEV))
    loss_o.backward()
    assert float(bpp_o.detach()) == pytest.approx(float(bpp_r.detach()), rel=2e-5)
    assert float(loss_o) == pytest.approx(float(loss_r), rel=2e-5)

    def close(name, a, r):
        a = a.detach().cpu()
        scale = max(r.abs().max().item(), 1e-6)
        err = (a - r).abs().max().item()
        assert err <= 2e-3 * scale, f"{name}: grad err {err:.3g} vs scale {scale:.3g}"

    for (n1, p1), (n2, p2) in zip(sorted(ours.named_parameters()), sorted(ref.named_parameters())):
        if n1.startswith(("entropy_bottleneck", "gaussian_conditional")):
            continue
    ref_named = dict(ref.named_parameters())
    for name, p in ours.named_parameters():
        if name in ref_named:
            close(name, p.grad, ref_named[name].grad)
    for i in range(5):
        close(f"_matrix{i}", getattr(ours.entropy_bottleneck, f"_matrix{i}").grad, ref.ref_eb.matrices[i].grad)
        close(f"_bias{i}", getattr(ours.entropy_bottleneck, f"_bias{i}").grad, ref.ref_eb.biases[i].grad)
        if i < 4:
            close(f"_factor{i}", getattr(ours.entropy_bottleneck, f"_factor{i}").grad, ref.ref_eb.factors[i].grad)
