"""CPU: the oracle restatement against the golden vectors generated from the reference's own
code (oracle/gen_golden.py).  Bit-exact everywhere: same machine class, same torch ops."""
import math

import pytest
import torch

from oracle import compressai_ref as cr
from tests.util import assert_equal_exact, load_golden


@pytest.fixture(scope="module")
def gc():
    return load_golden("gc_golden.npz")


@pytest.fixture(scope="module")
def eb():
    return load_golden("eb_golden.npz")


def test_scale_table_matches_reference(gc):
    assert_equal_exact(cr.get_scale_table(), gc["scale_table"], "scale table (tcm.py:33-34)")


@pytest.mark.parametrize("tag", ["", "edge_"])
def test_ste_round_and_likelihood_twin(gc, tag):
    y, mu, sg = gc[tag + "y"], gc[tag + "mu"], gc[tag + "sigma"]
    ste = cr.ste_round(y - mu) + mu
    assert_equal_exact(ste, gc[tag + "ste"], "ste_round(y-mu)+mu (tcm.py:457)")
    assert_equal_exact(cr.quantize(y, "dequantize", mu), gc[tag + "ste"], "quantize dequantize == ste_round")
    lik = cr.gc_likelihood(ste, sg, mu)
    assert_equal_exact(lik, gc[tag + "lik_unbounded"], "GC likelihood (tcm.py:570-582)")


def test_noise_mode_likelihood(gc):
    yhat, lik = cr.gc_forward(gc["y"], gc["sigma"], gc["mu"], training=True, noise=gc["noise"], likelihood_bound=0.0)
    assert_equal_exact(yhat, gc["y"] + gc["noise"], "noise quantize")
    assert_equal_exact(lik, gc["lik_noise_unbounded"], "noise-mode likelihood")


@pytest.mark.parametrize("tag", ["", "edge_"])
def test_build_indexes(gc, tag):
    idx = cr.build_indexes(gc[tag + "sigma"], gc["scale_table"])
    assert_equal_exact(idx, gc[tag + "indexes"], "build_indexes (adaptive_gaussian_conditional.py:606-617)")
    assert idx.dtype == torch.int32 and int(idx.min()) >= 0 and int(idx.max()) <= 63


def test_forward_equals_reference_stanh_module_default_params(gc):
    """The reference's own GaussianConditionalStanh (w=1, b=k+0.5) in eval mode coincides with
    the CompressAI-style path inside +-extrema and away from ties."""
    yhat, lik = cr.gc_forward(gc["y"], gc["sigma"], gc["mu"], training=False)
    m = gc["stanh_valid"]
    assert int(m.sum()) > 0.5 * m.numel()
    assert_equal_exact(yhat[m], gc["stanh_yhat"][m], "y_hat vs reference STanH module")
    assert_equal_exact(lik[m], gc["stanh_lik"][m], "bounded likelihood vs reference STanH module")


def test_entropy_bottleneck_vs_reference_module(eb):
    C = eb["z"].shape[1]
    ref = cr.EntropyBottleneckRef(C)
    ref.matrices = [eb[f"_matrix{i}"] for i in range(5)]
    ref.biases = [eb[f"_bias{i}"] for i in range(5)]
    ref.factors = [eb[f"_factor{i}"] for i in range(4)]
    zhat, lik = ref.forward(eb["z"], training=False)   # medians are 0 at init
    assert_equal_exact(zhat, eb["zhat"], "z_hat vs reference EntropyBottleneckStanh")
    assert_equal_exact(lik, eb["lik"], "z likelihood vs reference EntropyBottleneckStanh")


def test_pmf_to_quantized_cdf_properties():
    g = torch.Generator().manual_seed(5)
    for n in (1, 2, 3, 17, 64, 300):
        p = torch.rand(n, generator=g) ** 4
        p = p / p.sum()
        cdf = cr.pmf_to_quantized_cdf(p.tolist(), 16)
        assert cdf[0] == 0 and cdf[-1] == 1 << 16 and len(cdf) == n + 1
        assert all(b > a for a, b in zip(cdf, cdf[1:]))
    with pytest.raises(ValueError):
        cr.pmf_to_quantized_cdf([0.5, -0.1], 16)
    with pytest.raises(ValueError):
        cr.pmf_to_quantized_cdf([0.5, float("nan")], 16)


def test_gc_update_tables_shape_and_monotone():
    cdf, offset, length = cr.gc_update(cr.get_scale_table())
    assert cdf.shape[0] == 64 and offset.shape == (64,) and length.shape == (64,)
    for i in range(64):
        row = cdf[i, : int(length[i])]
        assert int(row[0]) == 0 and int(row[-1]) == 65536
        assert bool((row[1:] > row[:-1]).all())
    assert bool((offset[1:] <= offset[:-1]).all())


def test_bpp_matches_loss_formula():
    g = torch.Generator().manual_seed(1)
    l1 = torch.rand(2, 8, 4, 4, generator=g).clamp_min(1e-9)
    l2 = torch.rand(2, 3, 1, 1, generator=g).clamp_min(1e-9)
    n = 2 * 64 * 64
    ref = sum(torch.log(l).sum() / (-math.log(2) * n) for l in (l1, l2))   # loss.py:24-27
    assert float(cr.bpp([l1, l2], n)) == pytest.approx(float(ref), rel=0, abs=0)
    per = cr.per_image_bits(l1) + cr.per_image_bits(l2)
    assert float(per.sum() / n) == pytest.approx(float(ref), rel=1e-6)
