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


@pytest.mark.parametrize("shape", [(2, 16, 8, 8), (3, 5, 7, 3)])   # 128-bit kernel / scalar kernel (n % 4 != 0)
@pytest.mark.parametrize("training", [True, False])
def test_gc_backward_matches_autograd(training, shape):
    g = torch.Generator().manual_seed(17 + int(training))
    mu = torch.randn(shape, generator=g)
    sigma = torch.exp(torch.empty(shape).uniform_(-3.0, 3.0, generator=g))   # some below the 0.11 bound
    y = mu + sigma * torch.randn(shape, generator=g) * 1.5
    y.view(-1)[:4] = mu.view(-1)[:4] + torch.tensor([40.0, -40.0, 0.0, 0.5])  # floor-bounded likelihoods, |.|'(0)
    noise = torch.empty(shape).uniform_(-0.5, 0.5, generator=g)
    wy, wl, ws = (torch.randn(shape, generator=g) for _ in range(3))

    leaves = [t.clone().requires_grad_(True) for t in (y, sigma, mu)]
    out, lik, ste = _ref_forward(leaves[0], leaves[1], leaves[2], noise, training)
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


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("shape", [(3, 6, 5, 4), (40, 6, 16, 16)])   # one CTA per channel / channels cut over 8 CTAs
def test_eb_backward_matches_autograd(training, shape):
    from reslic_tcm_b200 import EntropyBottleneck, synthetic

    C = shape[1]
    params = synthetic.eb_parameters(C, trained_like=True, seed=99)
    g = torch.Generator().manual_seed(5)
    z = 2.5 * torch.randn(shape, generator=g)
    z.view(-1)[:3] = torch.tensor([60.0, -60.0, 0.0])          # likelihood floor on the first two
    noise = torch.empty(shape).uniform_(-0.5, 0.5, generator=g)
    wz, wl = torch.randn(shape, generator=g), torch.randn(shape, generator=g)

    # reference: the oracle's op sequence under torch autograd, with the LowerBound gradient rule
    ref = cr.EntropyBottleneckRef(C)
    ref.matrices = [params[f"_matrix{i}"].clone().requires_grad_(True) for i in range(5)]
    ref.biases = [params[f"_bias{i}"].clone().requires_grad_(True) for i in range(5)]
    ref.factors = [params[f"_factor{i}"].clone().requires_grad_(True) for i in range(4)]
    ref.quantiles = params["quantiles"].clone().requires_grad_(True)
    zr = z.clone().requires_grad_(True)
    xp = zr.permute(1, 0, 2, 3).reshape(C, 1, -1)
    if training:
        outputs = xp + noise.permute(1, 0, 2, 3).reshape(C, 1, -1)
    else:
        med = ref._get_medians()
        outputs = torch.round(xp - med) + med
    lik = RefLowerBound(1e-9)(ref._likelihood(outputs))
    back = lambda t: t.reshape(C, shape[0], shape[2], shape[3]).permute(1, 0, 2, 3)
    loss = (back(outputs) * wz).sum() + (torch.log(back(lik)) * wl).sum()
    loss.backward()

    mod = EntropyBottleneck(C).to(DEV).train(training)
    synthetic.load_eb_parameters(mod, params)
    zd = z.clone().to(DEV).requires_grad_(True)
    zh, lk = mod(zd, training=training, noise=noise.to(DEV) if training else None)
    ((zh * wz.to(DEV)).sum() + (torch.log(lk) * wl.to(DEV)).sum()).backward()

    def check(name, a, r):
        a, r = a.detach().cpu(), (torch.zeros_like(a.cpu()) if r is None else r)
        scale = max(r.abs().max().item(), 1e-3)
        err = (a - r).abs().max().item()
        assert err <= 3e-4 * scale, f"{name}: max err {err:.3g} vs scale {scale:.3g}"

    check("d/dz", zd.grad, zr.grad)
    for i in range(5):
        check(f"d/d_matrix{i}", getattr(mod, f"_matrix{i}").grad, ref.matrices[i].grad)
        check(f"d/d_bias{i}", getattr(mod, f"_bias{i}").grad, ref.biases[i].grad)
        if i < 4:
            check(f"d/d_factor{i}", getattr(mod, f"_factor{i}").grad, ref.factors[i].grad)
    if not training:
        check("d/dquantiles", mod.quantiles.grad, ref.quantiles.grad)


def test_eb_backward_split_channels_are_bit_reproducible_and_leave_the_workspace_clean():
    from reslic_tcm_b200 import EntropyBottleneck, ops, synthetic

    C = 12
    mod = EntropyBottleneck(C).to(DEV).train()
    synthetic.load_eb_parameters(mod, synthetic.eb_parameters(C, trained_like=True, seed=3))
    m, b, f = mod._params()
    med = mod._medians_flat()
    torch.manual_seed(4)
    z = 3.0 * torch.randn(64, C, 8, 8, device=DEV)
    gz, gl = torch.randn_like(z), torch.randn_like(z)
    runs = [ops.eb_backward(z, m, b, f, med, training=True, g_zhat=gz, g_lik=gl, seed=9, offset=2) for _ in range(3)]
    for r in runs[1:]:
        assert torch.equal(r[0], runs[0][0])
        for a, c in zip(r[1] + r[2] + r[3], runs[0][1] + runs[0][2] + runs[0][3]):
            assert torch.equal(a, c)
    ws = ops._eb_bwd_workspace(z.device, C)
    assert int(ws[: 4 * C].view(torch.int32).abs().sum()) == 0          # arrival counters back to zero


@pytest.mark.parametrize("shape", [(300, 24, 4, 4), (37, 12, 2, 4), (9, 12, 3, 3)])
def test_eb_backward_regenerates_the_forward_noise_field(shape):
    """In-kernel Philox: the backward launch (quad form for hw % 4 == 0, per element otherwise) must see the very noise
    the forward launch of the same (seed, offset) added — checked against a backward launch that is GIVEN that noise."""
    from reslic_tcm_b200 import EntropyBottleneck, ops, synthetic

    C = shape[1]
    mod = EntropyBottleneck(C).to(DEV).train()
    synthetic.load_eb_parameters(mod, synthetic.eb_parameters(C, trained_like=True, seed=8))
    m, b, f = mod._params()
    med = mod._medians_flat()
    gen = torch.Generator(device=DEV).manual_seed(2)
    z = torch.randn(shape, device=DEV, generator=gen).mul_(4.0).round_().div_(4.0)      # quarter steps: z + u - z == u exactly
    gz, gl = torch.randn(shape, device=DEV, generator=gen), torch.randn(shape, device=DEV, generator=gen)
    fwd = ops.eb_forward(z, m, b, f, med, training=True, want=("zhat",), seed=31, offset=7)
    u = fwd.zhat - z
    assert float(u.abs().max()) <= 0.5
    own = ops.eb_backward(z, m, b, f, med, training=True, g_zhat=gz, g_lik=gl, seed=31, offset=7)
    given = ops.eb_backward(z, m, b, f, med, training=True, noise=u, g_zhat=gz, g_lik=gl)
    flat = lambda r: [r[0]] + list(r[1]) + list(r[2]) + list(r[3])
    for a, c in zip(flat(own), flat(given)):
        scale = max(float(c.abs().max()), 1e-6)
        assert float((a - c).abs().max()) <= 2e-5 * scale
    other = ops.eb_backward(z, m, b, f, med, training=True, g_zhat=gz, g_lik=gl, seed=31, offset=8)
    assert not torch.equal(other[0], own[0])


@pytest.mark.parametrize("training,removing_mean,symmetry", [(True, True, False), (True, False, False),
                                                             (False, True, False), (True, True, True)])
def test_stanh_backward_matches_autograd(training, removing_mean, symmetry):
    from oracle import stanh_ref as sr
    from reslic_tcm_b200 import stanh

    beta, extrema = 3.0, 6
    cfg = dict(beta=beta, num_sigmoids=0, extrema=extrema, trainable=False, removing_mean=removing_mean,
               symmetry=symmetry)
    mod = stanh.GaussianConditionalStanh(None, channels=4, gaussian_configuration=cfg).to(DEV)
    g = torch.Generator().manual_seed(41)
    with torch.no_grad():
        mod.stanh.w.mul_((1.0 + 0.2 * torch.rand(mod.stanh.w.shape, generator=g)).to(DEV))
    mod.stanh.update_state(torch.device(DEV))
    st = mod.stanh
    w = (st.sym_w if symmetry else st.w).detach().cpu()
    b = torch.sort((st.sym_b if symmetry else st.b).detach().cpu())[0]
    cum_w = st.cum_w.cpu()
    shape = (2, 4, 5, 6)
    mu = torch.randn(shape, generator=g)
    sigma = torch.exp(torch.empty(shape).uniform_(-3.0, 2.0, generator=g))
    y = mu + 2.0 * torch.randn(shape, generator=g)
    wy, wl = torch.randn(shape, generator=g), torch.randn(shape, generator=g)

    leaves = [t.clone().requires_grad_(True) for t in (y, sigma, mu)]
    avg, dist = sr.mid_and_half_gaps(cum_w)
    yh = sr.quantize(leaves[0], "training" if training else "dequantize", leaves[2], w, b, beta, symmetry, removing_mean)
    values = yh - leaves[2]
    low, up = sr.define_v0_and_v1(values.detach(), avg, dist)
    s = RefLowerBound(0.11)(leaves[1])
    upper = cr.standardized_cumulative((low - values) / s) * (values >= 0) + \
        cr.standardized_cumulative((values + up) / s) * (values < 0)
    lower = cr.standardized_cumulative((-up - values) / s) * (values >= 0) + \
        cr.standardized_cumulative((values - low) / s) * (values < 0)
    lik = RefLowerBound(1e-9)(upper - lower)
    ((yh * wy).sum() + (torch.log(lik) * wl).sum()).backward()

    dl = [t.clone().to(DEV).requires_grad_(True) for t in (y, sigma, mu)]
    yh2, lik2 = mod(dl[0], dl[1], training=training, means=dl[2])
    ((yh2 * wy.to(DEV)).sum() + (torch.log(lik2) * wl.to(DEV)).sum()).backward()
    for name, a, r in zip(("d/dy", "d/dsigma", "d/dmu"), (t.grad for t in dl), (t.grad for t in leaves)):
        r = torch.zeros(shape) if r is None else r
        a = torch.zeros(shape) if a is None else a.cpu()
        scale = max(r.abs().max().item(), 1.0)
        err = (a - r).abs()
        assert bool((err <= 5e-4 * r.abs() + 2e-5 * scale).all()), f"{name}: max err {err.max():.3g} (scale {scale:.3g})"

@pytest.mark.parametrize("training,symmetry,num_sigmoids", [(True, False, 0), (True, True, 0), (False, False, 0),
                                                            (True, False, 6)])
def test_stanh_parameter_gradients_match_autograd(training, symmetry, num_sigmoids):
    """trainable=True (the reference default): d loss / d stanh.w and d stanh.b through the soft quantizer
    (dense [1,K,N] sum in the reference) and through distance_points inside _likelihood, against fp64 torch
    autograd on the oracle restatement built from leaf w / b exactly as the reference's update_state does."""
    from oracle import stanh_ref as sr
    from reslic_tcm_b200 import stanh

    beta, extrema = 2.5, 5
    cfg = dict(beta=beta, num_sigmoids=num_sigmoids, extrema=extrema, trainable=True, removing_mean=True,
               symmetry=symmetry)
    mod = stanh.GaussianConditionalStanh(None, channels=4, gaussian_configuration=cfg).to(DEV)
    g = torch.Generator().manual_seed(43)
    with torch.no_grad():
        mod.stanh.w.mul_((1.0 + 0.2 * torch.rand(mod.stanh.w.shape, generator=g)).to(DEV))
        mod.stanh.b.add_((0.1 * (torch.rand(mod.stanh.b.shape, generator=g) - 0.5)).to(DEV))
    mod.stanh.update_state(torch.device(DEV))
    assert mod.stanh.w.requires_grad and mod.stanh.b.requires_grad
    shape = (3, 4, 9, 7)
    mu = torch.randn(shape, generator=g)
    sigma = torch.exp(torch.empty(shape).uniform_(-2.0, 1.5, generator=g))
    y = mu + 2.5 * torch.randn(shape, generator=g)
    wy, wl = torch.randn(shape, generator=g), torch.randn(shape, generator=g)

    # ---- reference: fp64 autograd from leaf w, b
    w_leaf = mod.stanh.w.detach().cpu().double().requires_grad_(True)
    b_leaf = mod.stanh.b.detach().cpu().double().requires_grad_(True)
    if symmetry:
        w_eff = torch.cat((torch.flip(w_leaf, [0]), w_leaf), 0)
        b_eff = torch.sort(torch.cat((torch.flip(-b_leaf, [0]), b_leaf), 0))[0]
        half = torch.cat((torch.zeros(1, dtype=torch.float64), torch.cumsum(w_leaf, 0)))
        cum_w = torch.cat((-torch.flip(half[1:], dims=[0]), half), dim=0)
    else:
        w_eff, b_eff = w_leaf, torch.sort(b_leaf)[0]
        cum_w = torch.cat((torch.zeros(1, dtype=torch.float64), torch.cumsum(w_leaf, 0))) - w_leaf.sum().detach() / 2
    avg, dist = sr.mid_and_half_gaps(cum_w)
    yd, sd, md = y.double(), sigma.double(), mu.double()
    yh = sr.quantize(yd, "training" if training else "dequantize", md, w_eff, b_eff, beta, symmetry, True)
    values = yh - md
    # one-hot cell selection in the fp32 tables the kernel uses (cell membership carries no gradient)
    j = torch.bucketize(values.detach().float(), mod.stanh.average_points.detach().cpu().float(), right=False)
    dl = torch.cat((torch.zeros(1, dtype=torch.float64), dist))
    dr = torch.cat((dist, torch.zeros(1, dtype=torch.float64)))
    low, up = dl[j], dr[j]
    s = torch.clamp(sd, min=0.11)
    upper = cr.standardized_cumulative((low - values) / s) * (values >= 0) + \
        cr.standardized_cumulative((values + up) / s) * (values < 0)
    lower = cr.standardized_cumulative((-up - values) / s) * (values >= 0) + \
        cr.standardized_cumulative((values - low) / s) * (values < 0)
    lik = RefLowerBound(1e-9)(upper - lower)      # compressai's rule: also passes gradients that push L up
    ((yh * wy.double()).sum() + (torch.log(lik) * wl.double()).sum()).backward()

    # ---- ours
    yh2, lik2 = mod(y.to(DEV), sigma.to(DEV), training=training, means=mu.to(DEV))
    assert yh2.requires_grad or lik2.requires_grad
    ((yh2 * wy.to(DEV)).sum() + (torch.log(lik2) * wl.to(DEV)).sum()).backward()
    for name, a, r in (("d/dw", mod.stanh.w.grad, w_leaf.grad), ("d/db", mod.stanh.b.grad, b_leaf.grad)):
        r = torch.zeros_like(w_leaf) if r is None else r
        a = torch.zeros_like(r) if a is None else a.detach().cpu().double()
        scale = max(r.abs().max().item(), 1.0)
        err = (a - r).abs()
        assert bool((err <= 2e-3 * r.abs() + 2e-4 * scale).all()), f"{name}: max err {err.max():.3g} (scale {scale:.3g}) ours {a} ref {r}"
    if training:
        assert float(mod.stanh.b.grad.abs().max()) > 0 and float(mod.stanh.w.grad.abs().max()) > 0


def test_lrp_tail_in_place_on_a_slice():
    from reslic_tcm_b200 import ops

    g = torch.Generator().manual_seed(2)
    full = torch.randn(3, 320, 4, 4, generator=g).to(DEV)
    lrp = torch.randn(3, 64, 4, 4, generator=g).to(DEV)
    ref = full.clone()
    ref[:, 64:128] += 0.5 * torch.tanh(lrp)
    ops.lrp_tail_(full[:, 64:128], lrp)
    assert torch.allclose(full, ref, rtol=0, atol=2e-7)
