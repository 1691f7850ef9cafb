import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stanh_ref as sr, compressai_ref as cr
from reslic_tcm_b200 import stanh
DEV = "cuda:0"
beta, extrema = 2.5, 5
cfg = dict(beta=beta, num_sigmoids=0, extrema=extrema, trainable=True, removing_mean=True, symmetry=False)
mod = stanh.GaussianConditionalStanh(None, channels=4, gaussian_configuration=cfg).to(DEV)
g = torch.Generator().manual_seed(43)
with torch.no_grad():
    mod.stanh.w.mul_((1.0 + 0.2 * torch.rand(mod.stanh.w.shape, generator=g)).to(DEV))
mod.stanh.update_state(torch.device(DEV))
shape = (3, 4, 9, 7)
mu = torch.randn(shape, generator=g)
sigma = torch.exp(torch.empty(shape).uniform_(-2.0, 1.5, generator=g))
y = mu + 2.5 * torch.randn(shape, generator=g)
for which in ("yh", "lik"):
    wy = torch.randn(shape, generator=torch.Generator().manual_seed(1)) if which == "yh" else torch.zeros(shape)
    wl = torch.randn(shape, generator=torch.Generator().manual_seed(2)) if which == "lik" else torch.zeros(shape)
    w_leaf = mod.stanh.w.detach().cpu().double().requires_grad_(True)
    b_leaf = mod.stanh.b.detach().cpu().double().requires_grad_(True)
    w_eff, b_eff = w_leaf, torch.sort(b_leaf)[0]
    cum_w = torch.cat((torch.zeros(1, dtype=torch.float64), torch.cumsum(w_leaf, 0))) - w_leaf.sum().detach() / 2
    avg, dist = sr.mid_and_half_gaps(cum_w)
    yd, sd, md = y.double(), sigma.double(), mu.double()
    yh = sr.quantize(yd, "training", md, w_eff, b_eff, beta, False, True)
    values = yh - md
    j = torch.bucketize(values.detach().float(), mod.stanh.average_points.detach().cpu().float(), right=False)
    dl = torch.cat((torch.zeros(1, dtype=torch.float64), dist)); dr = torch.cat((dist, torch.zeros(1, dtype=torch.float64)))
    for detach_dist in (False, True):
        for t in (w_leaf, b_leaf):
            t.grad = None
        low, up = dl[j], dr[j]
        if detach_dist:
            low, up = low.detach(), up.detach()
        s = torch.clamp(sd, min=0.11)
        upper = cr.standardized_cumulative((low - values) / s) * (values >= 0) + cr.standardized_cumulative((values + up) / s) * (values < 0)
        lower = cr.standardized_cumulative((-up - values) / s) * (values >= 0) + cr.standardized_cumulative((values - low) / s) * (values < 0)
        lik = torch.clamp(upper - lower, min=1e-9)
        ((yh * wy.double()).sum() + (torch.log(lik) * wl.double()).sum()).backward(retain_graph=True)
        print(which, "ref detach_dist", detach_dist, "gw", w_leaf.grad.numpy().round(3), "gb", b_leaf.grad.numpy().round(3))
    mod.stanh.w.grad = None; mod.stanh.b.grad = None
    yh2, lik2 = mod(y.to(DEV), sigma.to(DEV), training=True, means=mu.to(DEV))
    ((yh2 * wy.to(DEV)).sum() + (torch.log(lik2) * wl.to(DEV)).sum()).backward()
    print(which, "ours gw", mod.stanh.w.grad.cpu().numpy().round(3), "gb", mod.stanh.b.grad.cpu().numpy().round(3))
