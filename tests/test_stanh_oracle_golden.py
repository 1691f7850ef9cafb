"""CPU: the STanH oracle restatement against golden vectors produced by the reference's own
STanH modules (oracle/gen_golden.py:gen_stanh).  Bit-exact (same torch ops, same order)."""
import pytest
import torch

from oracle import stanh_ref as sr
from tests.util import assert_equal_exact, load_golden


@pytest.fixture(scope="module")
def g():
    return load_golden("stanh_golden.npz")


def _case(g, tag):
    sym, extrema, beta, rm = (float(v) for v in g[f"{tag}_meta"])
    return dict(symmetric=bool(sym), beta=beta, removing_mean=bool(rm), w=g[f"{tag}_w"], b=g[f"{tag}_b"],
                cum_w=g[f"{tag}_cum_w"], y=g[f"{tag}_y"], mu=g[f"{tag}_mu"], sigma=g[f"{tag}_sigma"])


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_levels_and_activation(g, tag):
    c = _case(g, tag)
    if c["symmetric"]:
        cum = sr.levels_sym(g[f"{tag}_w_param"])
    else:
        cum = sr.levels_nonsym(g[f"{tag}_w_param"])
    assert_equal_exact(cum, c["cum_w"], "cum_w")
    avg, dist = sr.mid_and_half_gaps(cum)
    assert_equal_exact(avg, g[f"{tag}_avg"], "average_points")
    assert_equal_exact(dist, g[f"{tag}_dist"], "distance_points")
    assert_equal_exact(sr.stanh(c["y"], c["w"], c["b"], -1, c["symmetric"]), g[f"{tag}_hard"], "hard STanH")
    assert_equal_exact(sr.stanh(c["y"], c["w"], c["b"], c["beta"], c["symmetric"]), g[f"{tag}_soft"], "soft STanH")
    assert_equal_exact(sr.gap(c["y"], c["w"], c["b"], c["beta"], c["symmetric"]).reshape(1), g[f"{tag}_gap"], "gap")


@pytest.mark.parametrize("tag", ["A", "B", "C"])
@pytest.mark.parametrize("training", [False, True])
def test_forward(g, tag, training):
    c = _case(g, tag)
    yh, lik = sr.forward(c["y"], c["sigma"], c["mu"], training, c["w"], c["b"], c["cum_w"], c["beta"],
                         c["symmetric"], c["removing_mean"])
    key = "train" if training else "eval"
    assert_equal_exact(yh, g[f"{tag}_yhat_{key}"], "y_hat")
    assert_equal_exact(lik, g[f"{tag}_lik_{key}"], "likelihood")


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_symbols_and_unbounded_likelihood(g, tag):
    c = _case(g, tag)
    assert_equal_exact(sr.symbols(c["y"], c["mu"], c["cum_w"], c["w"], c["b"], c["symmetric"]), g[f"{tag}_sym"], "symbols")
    avg, dist = sr.mid_and_half_gaps(c["cum_w"])
    lik = sr.likelihood(g[f"{tag}_yhat_train"], c["sigma"], c["mu"], avg, dist)
    assert_equal_exact(lik, g[f"{tag}_lik_unbounded_train"], "_likelihood on given values")


@pytest.mark.parametrize("tag", ["N", "S"])
@pytest.mark.parametrize("training", [False, True])
def test_eb_stanh_forward(tag, training):
    from oracle import compressai_ref as cr

    g = load_golden("eb_stanh_golden.npz")
    C = g[f"{tag}_z"].shape[1]
    eb = cr.EntropyBottleneckRef(C)
    eb.matrices = [g[f"{tag}__matrix{i}"] for i in range(5)]
    eb.biases = [g[f"{tag}__bias{i}"] for i in range(5)]
    eb.factors = [g[f"{tag}__factor{i}"] for i in range(4)]
    zh, lik = sr.eb_stanh_forward(g[f"{tag}_z"], eb, g[f"{tag}_w"], g[f"{tag}_b"], g[f"{tag}_cum_w"], 4.0, tag == "S",
                                  training)
    key = "train" if training else "eval"
    assert_equal_exact(zh, g[f"{tag}_zhat_{key}"], "z_hat vs reference EntropyBottleneckStanh")
    assert_equal_exact(lik, g[f"{tag}_lik_{key}"], "likelihood vs reference EntropyBottleneckStanh")
