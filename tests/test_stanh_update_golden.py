"""CPU: the CDF-table builders of the STanH entropy models against the reference's OWN update()
methods — GaussianConditionalStanh.update (src/entropy_models/adaptive_gaussian_conditional.py:397-454) and
EntropyBottleneckStanh.update (src/entropy_models/adaptive_entropy_bottleneck.py:481-514) — run unmodified under the
shim by oracle/gen_golden.py:gen_stanh_update and committed as tests/golden/stanh_update_golden.npz.  update() is
setup code (torch + the host C++ pmf_to_quantized_cdf), so it is checked here without a GPU."""
import pytest
import torch

from reslic_tcm_b200.stanh import EntropyBottleneckStanh, GaussianConditionalStanh
from tests.util import assert_equal_exact, load_golden

CPU = torch.device("cpu")


@pytest.fixture(scope="module")
def g():
    return load_golden("stanh_update_golden.npz")


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_gaussian_conditional_stanh_update_matches_the_reference(g, tag):
    sym, extrema, beta = (float(v) for v in g[f"gc{tag}_meta"])
    cfg = dict(beta=beta, num_sigmoids=0, extrema=int(extrema), trainable=True, removing_mean=True, symmetry=bool(sym))
    m = GaussianConditionalStanh(None, channels=4, gaussian_configuration=cfg)
    with torch.no_grad():
        m.stanh.w.copy_(g[f"gc{tag}_w_param"])
        m.stanh.b.copy_(g[f"gc{tag}_b_param"])
    m.scale_table = g[f"gc{tag}_scale_table"].clone()
    m.update(CPU)
    assert_equal_exact(m.stanh.cum_w, g[f"gc{tag}_cum_w"], "levels")
    assert_equal_exact(m.pmf, g[f"gc{tag}_pmf"], "pmf over the levels, per scale")
    assert_equal_exact(m.cdf, g[f"gc{tag}_cdf"], "float cdf")
    assert_equal_exact(m._cdf_length.reshape(-1), g[f"gc{tag}_cdf_length"], "_cdf_length")
    assert_equal_exact(m._quantized_cdf, g[f"gc{tag}_quantized_cdf"], "_quantized_cdf (row assembly of _pmf_to_cdf)")
    # every row is a valid coder table: starts at 0, ends at 2^16, strictly increasing over its length
    q = m._quantized_cdf.long()
    for i in range(q.shape[0]):
        n = int(m._cdf_length[i])
        assert q[i, 0] == 0 and q[i, n - 1] == 1 << 16 and bool((q[i, 1:n] > q[i, :n - 1]).all())
    # documented deviation: the reference stores -cum_w[0] (a float level) as _offset; the drop-in
    # stores the integer index of the first level, one per row, which is what symbol - offset needs
    assert float(g[f"gc{tag}_offset"][0]) == -float(g[f"gc{tag}_cum_w"][0])
    assert m._offset.dtype == torch.int32 and m._offset.shape == (q.shape[0],)
    assert int(m._offset[0]) == m.stanh.symbol_offset


@pytest.mark.parametrize("tag", ["N", "S"])
def test_entropy_bottleneck_stanh_update_matches_the_reference(g, tag):
    cfg = dict(beta=4, num_sigmoids=0, extrema=6, trainable=True, symmetry=(tag == "S"))
    m = EntropyBottleneckStanh(5, factorized_configuration=cfg)
    with torch.no_grad():
        m.stanh.w.copy_(g[f"eb{tag}_w_param"])
        m.stanh.b.copy_(g[f"eb{tag}_b_param"])
        for i in range(5):
            getattr(m, f"_matrix{i}").copy_(g[f"eb{tag}__matrix{i}"])
            getattr(m, f"_bias{i}").copy_(g[f"eb{tag}__bias{i}"])
            if i < 4:
                getattr(m, f"_factor{i}").copy_(g[f"eb{tag}__factor{i}"])
    assert m.update(CPU) is True
    assert_equal_exact(m.stanh.cum_w, g[f"eb{tag}_cum_w"], "levels")
    assert_equal_exact(m.pmf, g[f"eb{tag}_pmf"], "pmf over the levels, per channel")
    assert_equal_exact(m.cdf, g[f"eb{tag}_cdf"], "float cdf")
