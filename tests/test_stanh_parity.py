"""GPU parity: the fused STanH kernel (through the C ABI) against the STanH oracle and the
golden vectors produced by the reference's own modules.

Bit-exact: level indexes ("symbols"), hard levels whenever the level table is exact in fp32
(default unit weights).  With perturbed weights the reference forms each value as an fp32 sum
of K terms in torch's reduction order; the kernel reads the level from the module's cum_w, which
can differ in the last bits -> y_hat within 1e-6 * max(1, |value|) (~8 ulps).  Soft (finite beta)
values: the same bound.  Likelihood: evaluated by the oracle ON THE KERNEL'S OWN y_hat, 1e-5 relative."""
import pytest
import torch

from oracle import stanh_ref as sr
from reslic_tcm_b200 import stanh
from tests.util import assert_equal_exact, assert_lik_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
YTOL = 1e-6      # ~8 fp32 ulps, relative to max(1, |value|)


@pytest.fixture(scope="module")
def g():
    return load_golden("stanh_golden.npz")


def _module(g, tag):
    sym, extrema, beta, rm = (float(v) for v in g[f"{tag}_meta"])
    cfg = dict(beta=beta, num_sigmoids=0, extrema=int(extrema), trainable=True, removing_mean=bool(rm),
               symmetry=bool(sym))
    m = stanh.GaussianConditionalStanh(None, channels=8, gaussian_configuration=cfg).to(DEV)
    with torch.no_grad():
        m.stanh.w.copy_(g[f"{tag}_w_param"])
        m.stanh.b.copy_(g[f"{tag}_b_param"])
    m.stanh.update_state(torch.device(DEV))
    return m, dict(symmetric=bool(sym), beta=beta, removing_mean=bool(rm), w=g[f"{tag}_w"], b=g[f"{tag}_b"],
                   cum_w=g[f"{tag}_cum_w"], y=g[f"{tag}_y"], mu=g[f"{tag}_mu"], sigma=g[f"{tag}_sigma"])


def _soft_tol(c):
    """The reference forms a soft value as an fp32 sum of K terms whose partial sums reach sum(w)/2
    (80 for extrema=80), so its OWN result carries rounding of a few ulp(sum(w)/2) ~ 4e-6 even for
    small values; the kernel adds only the unsaturated window to exact prefix sums."""
    return max(2e-6, 6e-8 * float(c["w"].sum()))


def _close(a, b, tol, what):
    """|a - b| <= tol * max(1, |b|): `tol` absolute for values below 1, a few fp32 ulps above (the
    reference's K-term fp32 sums carry rounding of that size themselves)."""
    a, b = a.detach().cpu().double(), b.double()
    err = ((a - b).abs() / b.abs().clamp_min(1.0)).max().item()
    assert err <= tol, f"{what}: max scaled err {err:.3g} > {tol}"


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_module_state_matches_reference(g, tag):
    m, c = _module(g, tag)
    assert_equal_exact(m.stanh.cum_w, c["cum_w"], "cum_w")
    assert_equal_exact(m.stanh.average_points, g[f"{tag}_avg"], "average_points")
    assert_equal_exact(m.stanh.distance_points, g[f"{tag}_dist"], "distance_points")


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_activation_hard_soft_and_gap(g, tag):
    m, c = _module(g, tag)
    y = c["y"].to(DEV)
    with torch.no_grad():
        hard = m.stanh(y, -1)
        soft = m.stanh(y, c["beta"])
        gap = stanh.compute_gap(m.stanh, y, c["beta"])
    if tag == "A":
        assert_equal_exact(hard, g[f"{tag}_hard"], "hard STanH (unit weights: exact)")
    _close(hard, g[f"{tag}_hard"], _soft_tol(c), "hard STanH")
    _close(soft, g[f"{tag}_soft"], _soft_tol(c), "soft STanH")
    assert float(gap) == pytest.approx(float(g[f"{tag}_gap"]), rel=2e-4, abs=1e-7)


@pytest.mark.parametrize("tag", ["A", "B", "C"])
@pytest.mark.parametrize("training", [False, True])
def test_forward_vs_reference_and_oracle(g, tag, training):
    m, c = _module(g, tag)
    key = "train" if training else "eval"
    with torch.no_grad():
        yh, lik = m(c["y"].to(DEV), c["sigma"].to(DEV), training=training, means=c["mu"].to(DEV))
        r = m.forward_fused(c["y"].to(DEV), c["sigma"].to(DEV), training=training, means=c["mu"].to(DEV))
    _close(yh, g[f"{tag}_yhat_{key}"], _soft_tol(c), "y_hat vs reference module")
    if tag == "A" and not training:
        assert_equal_exact(yh, g[f"{tag}_yhat_{key}"], "y_hat eval (unit weights: exact)")
        assert_lik_close(lik, g[f"{tag}_lik_{key}"], what="likelihood vs reference module")
    # likelihood: the reference formula evaluated on the kernel's own y_hat
    avg, dist = sr.mid_and_half_gaps(c["cum_w"])
    lik_ref = cr_bound(sr.likelihood(yh.cpu(), c["sigma"], c["mu"], avg, dist))
    assert_lik_close(lik, lik_ref, what="likelihood vs oracle on the same y_hat")
    # and loosely against the reference's own output (its y_hat differs in the last bits, which
    # small sigmas amplify: d ln L / dv ~ v / sigma^2)
    big = c["sigma"] > 0.5
    assert_lik_close(lik.cpu()[big], g[f"{tag}_lik_{key}"][big], rtol=2e-4, what="likelihood vs reference (sigma > .5)")
    own = -(torch.log2(lik.double()).reshape(lik.shape[0], -1).sum(1))
    assert torch.allclose(r["bits"], own, rtol=2e-6)


def cr_bound(x):
    from oracle import compressai_ref as cr

    return cr.lower_bound(x, 1e-9)


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_symbols_dequantize_and_likelihood_only(g, tag):
    m, c = _module(g, tag)
    y, mu, sg = c["y"].to(DEV), c["mu"].to(DEV), c["sigma"].to(DEV)
    with torch.no_grad():
        sym = m.quantize(y, "symbols", means=mu)
        deq = m.quantize(y, "dequantize", means=mu)
        back = m.dequantize(sym, means=mu)
        lik = m._likelihood(g[f"{tag}_yhat_train"].to(DEV), sg, means=mu)
    assert_equal_exact(sym, g[f"{tag}_sym"], "symbols vs the reference's per-element loop")
    _close(back, deq.cpu(), 1e-6, "dequantize(symbols) == quantize(dequantize)")
    assert_lik_close(lik, g[f"{tag}_lik_unbounded_train"], what="_likelihood on given values")


def test_large_random_against_oracle_and_edge_values():
    cfg = dict(beta=10, num_sigmoids=0, extrema=80, trainable=False, removing_mean=True, symmetry=False)
    m = stanh.GaussianConditionalStanh(None, channels=64, gaussian_configuration=cfg).to(DEV)
    m.stanh.update_state(torch.device(DEV))
    gen = torch.Generator().manual_seed(3)
    shape = (2, 64, 16, 16)
    mu = torch.randn(shape, generator=gen)
    sigma = torch.exp(torch.empty(shape).uniform_(-3, 5, generator=gen))
    y = mu + sigma * torch.randn(shape, generator=gen)
    y.view(-1)[:8] = torch.tensor([1e4, -1e4, 79.6, -80.4, 0.5, -0.5, float("nan"), 1200.0])
    mu.view(-1)[:8] = 0.0
    w, b = m.stanh.w.detach().cpu(), torch.sort(m.stanh.b.detach().cpu())[0]
    for training in (False, True):
        with torch.no_grad():
            yh, lik = m(y.to(DEV), sigma.to(DEV), training=training, means=mu.to(DEV))
        yh_ref, lik_ref = sr.forward(y, sigma, mu, training, w, b, m.stanh.cum_w.cpu(), 10, False, True)
        a, r = yh.cpu(), yh_ref
        tol = 1e-5 if training else YTOL
        bad = ~(((a - r).abs() <= tol * r.abs().clamp_min(1.0)) | (torch.isnan(a) & torch.isnan(r)))
        idx = torch.nonzero(bad.reshape(-1))[:5].reshape(-1).tolist()
        assert not idx, [(i, y.reshape(-1)[i].item(), mu.reshape(-1)[i].item(), a.reshape(-1)[i].item(),
                          r.reshape(-1)[i].item()) for i in idx]
        avg, dist = sr.mid_and_half_gaps(m.stanh.cum_w.cpu())
        assert_lik_close(lik, cr_bound(sr.likelihood(a, sigma, mu, avg, dist)), what=f"training={training}")
    # default STanH (unit weights) == plain rounding inside the range: same symbols as round(y - mu)
    with torch.no_grad():
        sym = m.quantize(y.to(DEV), "symbols", means=mu.to(DEV)).cpu()
    d = (y - mu)
    inside = d.abs() < 79.0
    ties = ((d - torch.floor(d)) - 0.5).abs() < 1e-6
    ok = inside & ~ties
    assert torch.equal(sym[ok] - 80, torch.round(d[ok]).int())


@pytest.mark.parametrize("tag", ["N", "S"])
@pytest.mark.parametrize("training", [False, True])
def test_entropy_bottleneck_stanh(tag, training):
    from oracle import compressai_ref as cr

    g = load_golden("eb_stanh_golden.npz")
    C = g[f"{tag}_z"].shape[1]
    cfg = dict(beta=4, num_sigmoids=0, extrema=6, trainable=True, symmetry=(tag == "S"))
    mod = stanh.EntropyBottleneckStanh(C, factorized_configuration=cfg).to(DEV)
    ref = cr.EntropyBottleneckRef(C)
    with torch.no_grad():
        mod.stanh.w.copy_(g[f"{tag}_w_param"])
        mod.stanh.b.copy_(g[f"{tag}_b_param"])
        for i in range(5):
            getattr(mod, f"_matrix{i}").copy_(g[f"{tag}__matrix{i}"])
            getattr(mod, f"_bias{i}").copy_(g[f"{tag}__bias{i}"])
            if i < 4:
                getattr(mod, f"_factor{i}").copy_(g[f"{tag}__factor{i}"])
    mod.stanh.update_state(torch.device(DEV))
    ref.matrices = [g[f"{tag}__matrix{i}"] for i in range(5)]
    ref.biases = [g[f"{tag}__bias{i}"] for i in range(5)]
    ref.factors = [g[f"{tag}__factor{i}"] for i in range(4)]
    key = "train" if training else "eval"
    with torch.no_grad():
        zh, lik = mod(g[f"{tag}_z"].to(DEV), training=training)
        r = mod.forward_fused(g[f"{tag}_z"].to(DEV), training=training)
    c = dict(w=g[f"{tag}_w"])
    _close(zh, g[f"{tag}_zhat_{key}"], _soft_tol(c), "z_hat vs reference EntropyBottleneckStanh")
    # likelihood of the kernel's own z_hat through the oracle's MLP (the reference's z_hat differs in
    # the last bits; the cumulative logits amplify that in the tails)
    avg, dist = sr.mid_and_half_gaps(g[f"{tag}_cum_w"])
    x = zh.cpu().permute(1, 0, 2, 3).reshape(C, 1, -1)
    low, up = sr.define_v0_and_v1(x.reshape(-1), avg, dist)
    lower = ref._logits_cumulative((x.reshape(-1) - low).reshape(C, 1, -1))
    upper = ref._logits_cumulative((x.reshape(-1) + up).reshape(C, 1, -1))
    sign = -torch.sign(lower + upper)
    lik_ref = cr_bound(torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower)))
    lik_ref = lik_ref.reshape(C, zh.shape[0], zh.shape[2], zh.shape[3]).permute(1, 0, 2, 3)
    assert_lik_close(lik, lik_ref, what="likelihood vs oracle on the same z_hat")
    assert_lik_close(lik, g[f"{tag}_lik_{key}"], rtol=5e-4, what="likelihood vs reference module")
    own = -(torch.log2(lik.double()).reshape(lik.shape[0], -1).sum(1))
    assert torch.allclose(r["bits"], own, rtol=2e-6)


def test_rate_modes_across_kernels_share_one_workspace():
    """RESLIC_RATE_DEFERRED / RESLIC_RATE_COLLECT through the other kernels that emit a rate (STanH
    conditional, STanH bottleneck, plain bottleneck in table / per-launch-table / noise mode): launches of
    different kernels defer into ONE workspace and the last one collects; the total equals the sum of the
    launches' immediate results exactly (integer accumulation of the same per-warp fixed-point terms)."""
    from reslic_tcm_b200 import EntropyBottleneck, _cabi, ops, synthetic
    from reslic_tcm_b200 import stanh as st

    B = 3
    g = torch.Generator().manual_seed(9)
    ws = torch.zeros(int(_cabi.load().reslic_workspace_bytes(B)), dtype=torch.uint8, device=DEV)
    cfg = dict(beta=3.0, num_sigmoids=0, extrema=6, trainable=False, removing_mean=True, symmetry=False)
    gcs = st.GaussianConditionalStanh(None, channels=8, gaussian_configuration=cfg).to(DEV)
    gcs.stanh.update_state(torch.device(DEV))
    y = (3.0 * torch.randn(B, 8, 6, 5, generator=g)).to(DEV)
    sg = torch.exp(torch.empty(B, 8, 6, 5).uniform_(-2.0, 1.0, generator=g)).to(DEV)
    mu = torch.randn(B, 8, 6, 5, generator=g).to(DEV)
    ebs = st.EntropyBottleneckStanh(16, factorized_configuration=dict(beta=3.0, num_sigmoids=0, extrema=6,
                                                                       trainable=False, symmetry=False)).to(DEV)
    ebs.stanh.update_state(torch.device(DEV))
    eb = EntropyBottleneck(16).to(DEV).eval()
    synthetic.load_eb_parameters(eb, synthetic.eb_parameters(16))
    z = (2.0 * torch.randn(B, 16, 4, 3, generator=g)).to(DEV)
    m, b, f = eb._params()
    med = eb._medians_flat()

    def launches(mode):
        """mode(k, last) -> out dict for launch k."""
        outs = []
        with torch.no_grad():
            outs.append(gcs.forward_fused(y, sg, training=False, means=mu, want=("bits",), out=mode(0)))
            outs.append(ebs.forward_fused(z, training=False, want=("bits",), out=mode(1)))
            outs.append(ops.eb_forward(z, m, b, f, med, want=("bits",), out=mode(2)))                       # per-launch table
            outs.append(ops.eb_forward(z, m, b, f, med, want=("bits",), out=mode(3), lut=eb._eval_lut()))   # cached table
            outs.append(ops.eb_forward(z, m, b, f, med, training=True, want=("bits",), out=mode(4), seed=5))  # noise mode
        return outs

    imm = launches(lambda k: {})
    total = sum(o["bits"] if isinstance(o, dict) else o.bits for o in imm)
    final = torch.empty(B, dtype=torch.float64, device=DEV)
    res = launches(lambda k: {"workspace": ws, "bits_deferred": True} if k < 4
                   else {"workspace": ws, "bits": final, "bits_collect": True})
    assert all((o["bits"] if isinstance(o, dict) else o.bits) is None for o in res[:4])
    assert torch.equal(final, total)
    assert int(ws.view(torch.int64).abs().sum()) == 0
    # all deferred + finalize
    launches(lambda k: {"workspace": ws, "bits_deferred": True})
    assert torch.equal(ops.rate_finalize(ws, B), total)
    assert int(ws.view(torch.int64).abs().sum()) == 0


@pytest.mark.parametrize("extrema", [80, 200])         # K = 160 (KMAX 256 tables) / K = 400 (KMAX 1024 tables)
@pytest.mark.parametrize("beta", [3.0, -1])
def test_trained_like_nonuniform_tables_against_oracle(extrema, beta):
    """The cell-grid lookups must be exact for ANY ascending table: weights and thresholds perturbed as a trained
    module would have them, including two thresholds 1e-4 apart (several entries in one grid cell), inputs that sit
    exactly on thresholds and far outside the table."""
    cfg = dict(beta=beta, num_sigmoids=0, extrema=extrema, trainable=True, removing_mean=True, symmetry=False)
    m = stanh.GaussianConditionalStanh(None, channels=8, gaussian_configuration=cfg).to(DEV)
    gen = torch.Generator().manual_seed(11 + extrema)
    K = 2 * extrema
    with torch.no_grad():
        w = 1.0 + 0.3 * (2 * torch.rand(K, generator=gen) - 1)
        b = m.stanh.b.detach().cpu() + 0.3 * (2 * torch.rand(K, generator=gen) - 1)
        b = torch.sort(b)[0]
        b[K // 2 + 3] = b[K // 2 + 2] + 1e-4
        b[K // 2 + 4] = b[K // 2 + 2] + 2e-4
        b = torch.sort(b)[0]
        m.stanh.w.copy_(w)
        m.stanh.b.copy_(b)
    m.stanh.update_state(torch.device(DEV))
    cum_w = m.stanh.cum_w.detach().cpu()
    shape = (2, 8, 16, 16)
    mu = torch.randn(shape, generator=gen)
    sigma = torch.exp(torch.empty(shape).uniform_(-3, 4, generator=gen))
    y = mu + sigma * torch.randn(shape, generator=gen) * 2
    y.view(-1)[:K] = b + mu.view(-1)[:K]                  # exactly ON the thresholds (up to the rounding of the sum)
    mu.view(-1)[K:K + 6] = 0.0
    y.view(-1)[K:K + 6] = torch.tensor([b[K // 2 + 2], b[K // 2 + 3], b[K // 2 + 4], 3.0 * extrema, -3.0 * extrema, float("nan")])
    tol = max(2e-6, 6e-8 * float(w.sum())) * 4
    for training in (False, True):
        with torch.no_grad():
            yh, lik = m(y.to(DEV), sigma.to(DEV), training=training, means=mu.to(DEV))
        yh_ref = sr.quantize(y, "training" if training else "dequantize", mu, w, b, beta, False, True)
        a = yh.cpu()
        ok = ((a - yh_ref).abs() <= tol * yh_ref.abs().clamp_min(1.0)) | (torch.isnan(a) & torch.isnan(yh_ref))
        bad = torch.nonzero(~ok.reshape(-1))[:5].reshape(-1).tolist()
        assert not bad, [(i, y.reshape(-1)[i].item(), a.reshape(-1)[i].item(), yh_ref.reshape(-1)[i].item()) for i in bad]
        avg, dist = sr.mid_and_half_gaps(cum_w)
        assert_lik_close(lik, cr_bound(sr.likelihood(a, sigma, mu, avg, dist)), what=f"training={training}")
    with torch.no_grad():
        sym = m.quantize(y.to(DEV), "symbols", means=mu.to(DEV)).cpu()
    finite = ~torch.isnan(y)
    assert torch.equal(sym[finite], sr.symbols(y, mu, cum_w, w, b, False)[finite])
