"""GPU parity of the STanH model's entropy pass (pipeline.TcmStanhEntropyPath) against the oracle, restating
what src/models/stanh/tcm_stanh.py:396-451 runs between the dense transforms:

    z -> EntropyBottleneck (plain, :403)                      likelihoods, bits_z
    per slice: GaussianConditionalStanh(y_k, scale_k, means=mu_k, training=tr)   (:432)
               y_hat_k = ste_round(y_k - mu_k) + mu_k  (frozen STanH, :433-434)
    y_gap = quantize(y, "training"); gap = compute_gap(y, y_gap)                 (:448-449, 465-478)

Tolerances as tests/test_stanh_parity.py: ste values bit-exact, soft values 1e-5 scaled, likelihood evaluated by the
oracle ON the kernel's own quantizer output 1e-5 relative (+3e-7 absolute), bits 1e-5 relative, gap 2e-4 (a difference of
two fp32 means in the reference)."""
import pytest
import torch

from oracle import compressai_ref as cr
from oracle import stanh_ref as sr
from reslic_tcm_b200 import synthetic
from reslic_tcm_b200.pipeline import TcmStanhEntropyPath
from tests.util import assert_equal_exact, assert_lik_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(B, C, h, w, Cz, seed):
    gen = torch.Generator().manual_seed(seed)
    mu = torch.randn(B, C, h, w, generator=gen) * 2
    sigma = torch.exp(torch.empty(B, C, h, w).uniform_(-3.0, 4.0, generator=gen))
    y = mu + sigma * torch.randn(B, C, h, w, generator=gen)
    z = torch.randn(B, Cz, h // 4, w // 4, generator=gen) * 4
    noise_z = torch.empty_like(z).uniform_(-0.5, 0.5, generator=gen)
    return y, mu, sigma, z, noise_z


def _path(cfg, Cz):
    path = TcmStanhEntropyPath(cfg, z_channels=Cz, channels=8).to(DEV)
    params = synthetic.eb_parameters(channels=Cz)
    synthetic.load_eb_parameters(path.entropy_bottleneck, params)
    return path, params


def _eb_oracle(params, Cz, z, noise_z, training):
    ref = cr.EntropyBottleneckRef(Cz)
    ref.matrices = [params[f"_matrix{i}"] for i in range(5)]
    ref.biases = [params[f"_bias{i}"] for i in range(5)]
    ref.factors = [params[f"_factor{i}"] for i in range(4)]
    ref.quantiles = params["quantiles"]
    return ref.forward(z, training=training, noise=noise_z if training else None)[1]


@pytest.mark.parametrize("training,beta", [(True, 10.0), (True, 1.5), (False, 10.0)])
def test_stanh_step_matches_oracle(training, beta):
    B, C, h, w, Cz = 3, 20, 8, 12, 8
    cfg = dict(beta=beta, num_sigmoids=0, extrema=20, trainable=False, removing_mean=True, symmetry=False)
    path, params = _path(cfg, Cz)
    y, mu, sigma, z, noise_z = _inputs(B, C, h, w, Cz, 11)
    num_pixels = h * 16 * w * 16
    res = path(y.to(DEV), mu.to(DEV), sigma.to(DEV), z.to(DEV), training=training, num_pixels=num_pixels,
               noise_z=noise_z.to(DEV))
    st = path.gaussian_conditional[0].stanh
    wv, bv, cum_w = st.w.detach().cpu(), torch.sort(st.b.detach().cpu())[0], st.cum_w.cpu()
    avg, dist = sr.mid_and_half_gaps(cum_w)
    # the ste value every slice carries on
    assert_equal_exact(res["y_hat"], torch.round(y - mu) + mu, "ste_round(y - mu) + mu")
    # quantizer output and likelihood, slice by slice as the model calls them
    cs = C // path.num_slices
    bits_ref = torch.zeros(B, dtype=torch.float64)
    for k in range(path.num_slices):
        sl = slice(cs * k, cs * (k + 1))
        yq_ref, _ = sr.forward(y[:, sl], sigma[:, sl], mu[:, sl], training, wv, bv, cum_w, beta, False, True)
        a = res["y_q"][:, sl].cpu()
        tol = 1e-5 if training else 1e-6
        assert ((a - yq_ref).abs() <= tol * yq_ref.abs().clamp_min(1.0)).all(), f"slice {k}: quantizer output"
        lik_ref = cr.lower_bound(sr.likelihood(a, sigma[:, sl], mu[:, sl], avg, dist), 1e-9)
        assert_lik_close(res["likelihoods"]["y"][:, sl], lik_ref, what=f"slice {k} likelihood")
        bits_ref += -torch.log2(lik_ref.double()).sum(dim=(1, 2, 3))
    z_lik_ref = _eb_oracle(params, Cz, z, noise_z, training)
    assert_lik_close(res["likelihoods"]["z"], z_lik_ref, what="z likelihood")
    bits_ref += -torch.log2(z_lik_ref.double()).sum(dim=(1, 2, 3))
    got = res["bits"].cpu()
    assert ((got - bits_ref).abs() <= 1e-5 * bits_ref.abs()).all(), (got, bits_ref)
    # compute_gap over the whole y (no means), tcm_stanh.py:448-449
    gap = path.gap(res, y.numel()).item()
    gap_ref = sr.gap(y, wv, bv, beta, False).item()     # a difference of two fp32 means: 2e-4, as tests/test_stanh_parity.py
    assert gap == pytest.approx(gap_ref, rel=2e-4, abs=1e-7), (gap, gap_ref)


def test_stanh_step_graph_replay_and_module_calls_agree():
    """The captured pass equals the eager pass, and both equal the drop-in modules called one by one."""
    B, C, h, w, Cz = 2, 40, 8, 8, 8
    cfg = dict(beta=10.0, num_sigmoids=0, extrema=40, trainable=False, removing_mean=True, symmetry=False)
    path, _ = _path(cfg, Cz)
    y, mu, sigma, z, noise_z = (t.to(DEV) for t in _inputs(B, C, h, w, Cz, 5))
    kw = dict(training=True, num_pixels=h * w * 256, noise_z=noise_z)
    eager = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in path(y, mu, sigma, z, **kw).items() if k != "likelihoods"}
    lik_e = path._bufs["y_lik"].clone()
    graph, res = path.capture(y, mu, sigma, z, **kw)
    for t in (res["y_hat"], res["y_q"], res["bits"], res["gap_sums"], res["likelihoods"]["y"]):
        t.zero_()
    graph.replay()
    torch.cuda.synchronize()
    for k in ("y_hat", "y_q", "bits", "gap_sums"):
        assert torch.equal(res[k], eager[k]), k
    assert torch.equal(res["likelihoods"]["y"], lik_e)
    gc = path.gaussian_conditional[0]
    cs = C // path.num_slices
    for k in range(path.num_slices):
        sl = slice(cs * k, cs * (k + 1))
        with torch.no_grad():
            yq, lik = gc(y[:, sl], sigma[:, sl], means=mu[:, sl], training=True)
        assert torch.equal(yq, res["y_q"][:, sl]) and torch.equal(lik, res["likelihoods"]["y"][:, sl])
    from reslic_tcm_b200.stanh import compute_gap

    assert torch.equal(compute_gap(gc.stanh, y), path.gap(res, y.numel()))
