"""dev: launch ONE training-path kernel a few times on the config-5 slice shape (for ncu -k ... -s 2 -c 1).
usage: train_kernels_probe.py gap|stanh_soft|stanh_hard|stanh_bwd|eb_bwd|eb_fwd"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from reslic_tcm_b200 import EntropyBottleneck, ops, synthetic
from reslic_tcm_b200.stanh import GaussianConditionalStanh, compute_gap

op = sys.argv[1]
dev = torch.device("cuda:0")
B, C, h, w = 256, 64, 16, 16
g = torch.Generator(device=dev).manual_seed(1)
mu = torch.randn(B, C, h, w, device=dev, generator=g)
sigma = torch.exp(torch.empty(B, C, h, w, device=dev).uniform_(-3.0, 4.16, generator=g))
y = mu + sigma * torch.randn(B, C, h, w, device=dev, generator=g)
g1 = torch.randn(B, C, h, w, device=dev, generator=g)
g2 = torch.randn(B, C, h, w, device=dev, generator=g)
cfg = {"beta": 10.0, "num_sigmoids": 0, "extrema": 80, "symmetry": False, "trainable": False, "removing_mean": True}
m = GaussianConditionalStanh(None, channels=C, gaussian_configuration=cfg).to(dev)
m.stanh.update_state(dev)
Cz, hz = 192, 4
mod = EntropyBottleneck(Cz).to(dev).train()
synthetic.load_eb_parameters(mod, synthetic.eb_parameters())
mm, bb, ff = mod._params()
med = mod._medians_flat()
z = torch.randn(B, Cz, hz, hz, device=dev, generator=g) * 4
gz = torch.randn(B, Cz, hz, hz, device=dev, generator=g)
fns = {
    "gap": lambda: compute_gap(m.stanh, y),
    "stanh_soft": lambda: m.forward_fused(y, sigma, training=True, means=mu, want=("yhat", "lik")),
    "stanh_hard": lambda: m.forward_fused(y, sigma, training=False, means=mu, want=("yhat", "lik")),
    "stanh_bwd": lambda: m._stanh_backward(y, sigma, mu, True, g1, g2),
    "eb_bwd": lambda: ops.eb_backward(z, mm, bb, ff, med, training=True, g_zhat=gz, g_lik=gz, seed=1, offset=0),
    "eb_fwd": lambda: ops.eb_forward(z, mm, bb, ff, med, training=True, want=("zhat", "lik"), seed=1, offset=0),
}
for _ in range(4):
    fns[op]()
torch.cuda.synchronize()
print("ok", op)
