"""GPU parity: the fused Gaussian-conditional CUDA kernel (through the C ABI) against the CPU
oracle and the committed golden vectors.  Integers bit-exact; likelihood within
1e-5 relative (+2e-7 absolute, see tests/util.py); bits within 1e-5 relative."""
import math

import pytest
import torch

from oracle import compressai_ref as cr
from reslic_tcm_b200 import GaussianConditional, _cabi, ops, synthetic
from tests.util import LIK_ATOL_EXACT, assert_equal_exact, assert_lik_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def gc():
    m = GaussianConditional(None).to(DEV).eval()
    m.scale_table = cr.get_scale_table().to(DEV)
    return m


def _rand(shape, seed, sigma_hi=300.0):
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(shape, generator=g)
    sigma = torch.exp(torch.empty(shape).uniform_(math.log(0.05), math.log(sigma_hi), generator=g))
    y = mu + sigma * torch.randn(shape, generator=g)
    noise = torch.empty(shape).uniform_(-0.5, 0.5, generator=g)
    return y, mu, sigma, noise


def _check_all(gc, y, mu, sigma, training=False, noise=None, what=""):
    table = cr.get_scale_table()
    want = ("yhat", "ste", "lik", "sym", "idx", "bits")
    with torch.no_grad():
        r = gc.forward_fused(y.to(DEV), sigma.to(DEV), None if mu is None else mu.to(DEV), training=training,
                             want=want, noise=None if noise is None else noise.to(DEV))
    torch.cuda.synchronize()
    yhat_ref, lik_ref = cr.gc_forward(y, sigma, mu, training=training, noise=noise)
    m0 = mu if mu is not None else torch.zeros_like(y)
    assert_equal_exact(r.yhat, yhat_ref, what + " quantize output")
    assert_equal_exact(r.ste, cr.ste_round(y - m0) + m0, what + " ste_round output")
    assert_equal_exact(r.sym, cr.quantize(y, "symbols", mu), what + " symbols")
    assert_equal_exact(r.idx, cr.build_indexes(sigma, table), what + " indexes")
    assert_lik_close(r.lik, lik_ref, what=what + " likelihood")
    # vs the exact value of the same formula: fp64 evaluation on the fp32 operands
    # v = |y_hat - mu| and s the reference feeds into it
    v32 = (yhat_ref - m0).abs()
    lik64 = cr.lower_bound(cr.gc_likelihood(v32.double(), sigma.double(), None), 1e-9)
    assert_lik_close(r.lik, lik64.float(), atol=LIK_ATOL_EXACT, what=what + " likelihood vs fp64")
    bits_ref = cr.per_image_bits(lik_ref)
    assert torch.allclose(r.bits.cpu(), bits_ref, rtol=1e-5, atol=0), (what, r.bits.cpu(), bits_ref)
    # the kernel's own sum must agree with its own likelihood output much more tightly
    own = -(torch.log2(r.lik.double()).reshape(y.shape[0], -1).sum(1)).cpu()
    assert torch.allclose(r.bits.cpu(), own, rtol=2e-6, atol=1e-6)
    return r


@pytest.mark.parametrize("shape", [(1, 64, 16, 16), (3, 64, 12, 8), (2, 5, 3, 7), (1, 1, 1, 1), (4, 1, 1, 3),
                                   (2, 64, 48, 32)])
def test_eval_forward_all_outputs(gc, shape):
    y, mu, sigma, _ = _rand(shape, 100 + sum(shape))
    _check_all(gc, y, mu, sigma, what=f"eval {shape}")


@pytest.mark.parametrize("shape", [(1, 64, 16, 16), (2, 7, 5, 3), (2, 64, 24, 16)])
def test_noise_forward_explicit_noise(gc, shape):
    y, mu, sigma, noise = _rand(shape, 200 + sum(shape))
    _check_all(gc, y, mu, sigma, training=True, noise=noise, what=f"noise {shape}")


def test_no_means(gc):
    y, _, sigma, noise = _rand((2, 16, 8, 8), 7)
    _check_all(gc, y, None, sigma, what="means=None eval")
    _check_all(gc, y, None, sigma, training=True, noise=noise, what="means=None noise")


def test_channel_slices_of_full_latent_without_copy(gc):
    """tcm.py:438 y.chunk(5, 1): slices are strided views; outputs written into slices of
    preallocated full tensors."""
    B, h, w = 3, 6, 4
    y, mu, sigma, _ = _rand((B, 320, h, w), 9)
    yd, md, sd = y.to(DEV), mu.to(DEV), sigma.to(DEV)
    y_hat = torch.empty_like(yd)
    lik = torch.empty_like(yd)
    idx = torch.empty(yd.shape, dtype=torch.int32, device=DEV)
    bits = torch.zeros(B, dtype=torch.float64, device=DEV)
    with torch.no_grad():
        for k in range(5):
            sl = slice(64 * k, 64 * (k + 1))
            r = gc.forward_fused(yd[:, sl], sd[:, sl], md[:, sl], want=("ste", "lik", "idx", "bits"),
                                 out={"ste": y_hat[:, sl], "lik": lik[:, sl], "idx": idx[:, sl]})
            bits += r.bits
    yhat_ref, lik_ref = cr.gc_forward(y, sigma, mu)
    assert_equal_exact(y_hat, yhat_ref, "sliced y_hat")
    assert_equal_exact(idx, cr.build_indexes(sigma, cr.get_scale_table()), "sliced indexes")
    assert_lik_close(lik, lik_ref)
    assert torch.allclose(bits.cpu(), cr.per_image_bits(lik_ref), rtol=1e-5)


def test_misaligned_views_take_scalar_path(gc):
    y, mu, sigma, _ = _rand((2, 4, 5, 9), 21)
    yd = torch.zeros(2 * 4 * 5 * 9 + 1, device=DEV)[1:].view(2, 4, 5, 9)  # base pointer off by 4 bytes
    yd.copy_(y)
    with torch.no_grad():
        r = gc.forward_fused(yd, sigma.to(DEV), mu.to(DEV), want=("yhat", "lik", "sym"))
    yhat_ref, lik_ref = cr.gc_forward(y, sigma, mu)
    assert_equal_exact(r.yhat, yhat_ref)
    assert_equal_exact(r.sym, cr.quantize(y, "symbols", mu))
    assert_lik_close(r.lik, lik_ref)


def test_golden_vectors_from_reference_code(gc):
    g = load_golden("gc_golden.npz")
    for tag in ("", "edge_"):
        y, mu, sg = g[tag + "y"], g[tag + "mu"], g[tag + "sigma"]
        with torch.no_grad():
            r = gc.forward_fused(y.to(DEV), sg.to(DEV), mu.to(DEV), want=("ste", "idx", "sym"))
            lik = gc._likelihood(g[tag + "ste"].to(DEV), sg.to(DEV), mu.to(DEV))
        assert_equal_exact(r.ste, g[tag + "ste"], tag + "ste vs tcm.py ste_round")
        assert_equal_exact(r.idx, g[tag + "indexes"], tag + "indexes vs reference build_indexes")
        assert_lik_close(lik, g[tag + "lik_unbounded"], what=tag + "unbounded likelihood vs TCM._likelihood")
    with torch.no_grad():
        yh, lik = gc(g["y"].to(DEV), g["sigma"].to(DEV), g["mu"].to(DEV), training=False)
        ln = gc._likelihood((g["y"] + g["noise"]).to(DEV), g["sigma"].to(DEV), g["mu"].to(DEV))
    m = g["stanh_valid"]
    assert_equal_exact(yh.cpu()[m], g["stanh_yhat"][m], "forward vs reference STanH module")
    assert_lik_close(lik.cpu()[m], g["stanh_lik"][m])
    assert_lik_close(ln, g["lik_noise_unbounded"])


def test_edge_cases_ties_bounds_nan_inf(gc):
    table = cr.get_scale_table()
    nxt = lambda t, d: torch.nextafter(t, torch.full_like(t, d))
    sig = torch.cat([table, nxt(table, math.inf), nxt(table, -math.inf),
                     torch.tensor([0.0, -1.0, 1e-30, 0.10999, 0.11, 256.0, 257.0, 1e10, 3e38,
                                   float("inf"), float("nan")])])
    d = torch.tensor([-0.0, 0.0, 0.5, -0.5, 1.5, 2.5, -2.5, 0.49999997, 1e-8, 37.0, -113.0, 1000.0, 1e6, 3e9,
                      1e20, float("inf"), float("nan")])
    S, D = torch.meshgrid(sig, d, indexing="ij")
    mu = torch.full_like(S, 0.375)
    y = (D + mu).reshape(1, 1, *S.shape)
    mu, S = mu.reshape_as(y), S.reshape_as(y)
    with torch.no_grad():
        r = gc.forward_fused(y.to(DEV), S.to(DEV), mu.to(DEV), want=("yhat", "lik", "idx", "sym"))
    yhat_ref, lik_ref = cr.gc_forward(y, S, mu)
    assert_equal_exact(r.yhat, yhat_ref, "edge y_hat")
    assert_equal_exact(r.idx, cr.build_indexes(S, table), "edge indexes (NaN -> 63, ties at table values)")
    sane = torch.isfinite(y) & (y.abs() < 2e9)   # int32 conversion of out-of-range floats is undefined in torch
    assert_equal_exact(r.sym.cpu()[sane], cr.quantize(y, "symbols", mu)[sane], "edge symbols")
    # one documented deviation: sigma = +inf AND |y - mu| = +inf together give NaN in the
    # reference (inf/inf); the kernel clamps both to 1e30 and returns the likelihood bound.
    both_inf = torch.isinf(S) & torch.isinf(y - mu)
    lik = r.lik.cpu()
    assert bool((lik[both_inf] == 1e-9).all())
    lik_ref = torch.where(both_inf, torch.full_like(lik_ref, 1e-9), lik_ref)
    assert_lik_close(lik, lik_ref, what="edge likelihood")


def test_custom_scale_tables(gc):
    g = torch.Generator().manual_seed(3)
    sigma = torch.exp(torch.empty(2, 3, 17, 5).uniform_(-4, 7, generator=g))
    for n in (1, 2, 3, 17, 63, 64, 65, 200, 256):
        tab = torch.sort(torch.exp(torch.empty(n).uniform_(-2, 6, generator=g)))[0]
        got = ops.build_indexes(sigma.to(DEV), tab.to(DEV), 0.11)
        assert_equal_exact(got, cr.build_indexes(sigma, tab), f"table_len={n}")
    nan = torch.full((1, 1, 2, 2), float("nan"))
    for n in (1, 5, 64, 100):
        tab = torch.linspace(0.2, 9.0, n)
        assert_equal_exact(ops.build_indexes(nan.to(DEV), tab.to(DEV)), cr.build_indexes(nan, tab), f"NaN table_len={n}")


def test_quantize_modes_and_dequantize_roundtrip(gc):
    y, mu, _, _ = _rand((2, 8, 4, 6), 5)
    yd, md = y.to(DEV), mu.to(DEV)
    with torch.no_grad():
        sym = gc.quantize(yd, "symbols", md)
        deq = gc.quantize(yd, "dequantize", md)
        assert_equal_exact(sym, cr.quantize(y, "symbols", mu))
        assert_equal_exact(deq, cr.quantize(y, "dequantize", mu))
        assert_equal_exact(gc.dequantize(sym, md), cr.dequantize(cr.quantize(y, "symbols", mu), mu))
        assert_equal_exact(gc.dequantize(sym, md), deq, "dequantize(symbols) == quantize(dequantize)")
        assert_equal_exact(gc.quantize(yd, "symbols"), cr.quantize(y, "symbols"))
        nz = gc.quantize(yd, "noise")
    u = (nz - yd).cpu()
    assert float(u.abs().max()) <= 0.5 + 1e-6
    with pytest.raises(ValueError):
        gc.quantize(yd, "bogus")


def test_philox_noise_statistics_and_determinism(gc):
    y, mu, sigma, _ = _rand((4, 64, 32, 32), 11, sigma_hi=20.0)
    yd, md, sd = y.to(DEV), mu.to(DEV), sigma.to(DEV)
    a = ops.gc_forward(yd, sd, md, training=True, want=("yhat", "lik", "ste"), seed=42, offset=0)
    b = ops.gc_forward(yd, sd, md, training=True, want=("yhat",), seed=42, offset=0)
    c = ops.gc_forward(yd, sd, md, training=True, want=("yhat",), seed=42, offset=1)
    d = ops.gc_forward(yd, sd, md, training=True, want=("yhat",), seed=43, offset=0)
    assert torch.equal(a.yhat, b.yhat)
    assert not torch.equal(a.yhat, c.yhat) and not torch.equal(a.yhat, d.yhat)
    u = (a.yhat - yd).double().cpu().reshape(-1)
    keep = y.abs().reshape(-1) < 4.0          # where y + u keeps enough mantissa to recover u
    u = u[keep]
    assert float(u.abs().max()) < 0.5 + 1e-6
    n = u.numel()
    assert abs(float(u.mean())) < 5 * math.sqrt(1 / 12 / n)
    assert abs(float(u.var()) - 1 / 12) < 2e-3
    hist = torch.histc(u.float(), bins=16, min=-0.5, max=0.5) / n
    assert float((hist - 1 / 16).abs().max()) < 4e-3
    # likelihood is the reference formula applied to the noisy values the kernel returned
    lik_ref = cr.lower_bound(cr.gc_likelihood(a.yhat.cpu(), sigma, mu), 1e-9)
    assert_lik_close(a.lik, lik_ref, what="philox-mode likelihood")
    assert_equal_exact(a.ste, cr.ste_round(y - mu) + mu, "ste output in noise mode")


def test_empty_and_errors(gc):
    e = torch.empty(0, 64, 4, 4, device=DEV)
    r = ops.gc_forward(e, e, e, want=("yhat", "lik", "bits"))
    assert r.yhat.shape == e.shape and r.bits.shape == (0,)
    e2 = torch.empty(2, 0, 4, 4, device=DEV)
    r = ops.gc_forward(e2, e2, e2, want=("yhat", "lik"))
    assert r.lik.numel() == 0
    x = torch.zeros(1, 4, 2, 2, device=DEV)
    with pytest.raises(ValueError):
        ops.gc_forward(x, torch.zeros(1, 4, 2, 3, device=DEV), x)
    with pytest.raises(ValueError):
        GaussianConditional(None).to(DEV).build_indexes(x)
    with pytest.raises(TypeError):
        ops.gc_forward(x.double(), x.double(), x.double())
    d = _cabi.new(_cabi.GcDesc)
    assert _cabi.load().reslic_gc_fwd_f32(d, None) == 0          # B = 0: empty, ok
    d.B, d.n = 1, 4
    assert _cabi.load().reslic_gc_fwd_f32(d, None) == -1         # no pointers: argument error
    assert b"gc_fwd" in _cabi.load().reslic_last_error()


def test_workspace_reuse_across_batch_sizes(gc):
    """Regression: the cached workspace is shared by launches with different B; a counter slot
    must never alias an earlier launch's partial sums."""
    for B, C in [(1, 64), (8, 64), (3, 320), (24, 64), (2, 64), (16, 320), (5, 7)]:
        y, mu, sigma, _ = _rand((B, C, 16, 16), 31 + B)
        r = ops.gc_forward(y.to(DEV), sigma.to(DEV), mu.to(DEV), want=("lik", "bits"))
        own = -(torch.log2(r.lik.double()).reshape(B, -1).sum(1))
        assert torch.allclose(r.bits, own, rtol=2e-6), (B, C)

def test_deferred_rate_modes_equal_the_immediate_sum(gc):
    """RESLIC_RATE_DEFERRED launches leave fixed-point sums in the workspace; reslic_rate_finalize_f64 or a
    last RESLIC_RATE_COLLECT launch turns them into bits.  Integer accumulation: the total equals the sum
    of the per-launch immediate results exactly (both are multiples of 2^-16), the workspace ends zeroed,
    and non-finite partials (NaN input) survive the deferral."""
    B = 6
    parts = []
    ws = torch.zeros(int(_cabi.load().reslic_workspace_bytes(B)), dtype=torch.uint8, device=DEV)
    total = torch.zeros(B, dtype=torch.float64, device=DEV)
    for k in range(3):
        y, mu, sigma, _ = _rand((B, 64, 16, 16), 700 + k)
        parts.append((y.to(DEV), sigma.to(DEV), mu.to(DEV)))
        total += ops.gc_forward(*parts[-1], want=("bits",)).bits
    # (a) three deferred launches + finalize
    for a in parts:
        r = ops.gc_forward(*a, want=("lik", "bits"), out={"bits_deferred": True, "workspace": ws})
        assert r.bits is None
    got = ops.rate_finalize(ws, B)
    assert torch.equal(got, total)
    assert int(ws.view(torch.int64).abs().sum()) == 0
    # (b) accumulate into an existing vector
    for a in parts[:2]:
        ops.gc_forward(*a, want=("bits",), out={"bits_deferred": True, "workspace": ws})
    base = torch.full((B,), 10.0, dtype=torch.float64, device=DEV)
    ops.rate_finalize(ws, B, bits=base, accumulate=True)
    two = ops.gc_forward(*parts[0], want=("bits",)).bits + ops.gc_forward(*parts[1], want=("bits",)).bits
    assert torch.equal(base, two + 10.0)
    # (c) two deferred launches, the third collects
    for a in parts[:2]:
        ops.gc_forward(*a, want=("bits",), out={"bits_deferred": True, "workspace": ws})
    out_bits = torch.empty(B, dtype=torch.float64, device=DEV)
    r = ops.gc_forward(*parts[2], want=("bits",), out={"bits": out_bits, "bits_collect": True, "workspace": ws})
    assert torch.equal(r.bits, total)
    assert int(ws.view(torch.int64).abs().sum()) == 0
    # (d) a NaN in image 2 of a deferred launch is carried to the collected result
    y, sg, mu = parts[0]
    y_bad = y.clone(); y_bad[2, 3, 4, 5] = float("nan")
    ops.gc_forward(y_bad, sg, mu, want=("bits",), out={"bits_deferred": True, "workspace": ws})
    r = ops.gc_forward(*parts[1], want=("bits",), out={"bits_collect": True, "workspace": ws})
    assert torch.isnan(r.bits[2]) and torch.isfinite(r.bits[[0, 1, 3, 4, 5]]).all()
    assert int(ws.view(torch.int64).abs().sum()) == 0
    # (e) misuse
    with pytest.raises(ValueError):
        ops.gc_forward(*parts[0], want=("bits",), out={"bits_deferred": True})


def test_mirror_mode_matches_oracle_and_fast_mode(gc):
    """MIRROR arithmetic (CUDA erfcf/log2f = what the reference's torch CUDA kernels run) and
    the default FAST arithmetic agree within the likelihood tolerance; integers identical."""
    y, mu, sigma, _ = _rand((4, 64, 24, 16), 55)
    yd, md, sd = y.to(DEV), mu.to(DEV), sigma.to(DEV)
    table = cr.get_scale_table().to(DEV)
    want = ("ste", "lik", "sym", "idx", "bits")
    fast = ops.gc_forward(yd, sd, md, want=want, scale_table=table)
    _cabi.set_math_mode(_cabi.MATH_MIRROR)
    try:
        mirror = ops.gc_forward(yd, sd, md, want=want, scale_table=table)
    finally:
        _cabi.set_math_mode(_cabi.MATH_FAST)
    assert torch.equal(fast.ste, mirror.ste) and torch.equal(fast.sym, mirror.sym) and torch.equal(fast.idx, mirror.idx)
    _, lik_ref = cr.gc_forward(y, sigma, mu)
    assert_lik_close(mirror.lik, lik_ref, what="mirror vs oracle")
    assert_lik_close(fast.lik, mirror.lik.cpu(), what="fast vs mirror")
    assert torch.allclose(fast.bits, mirror.bits, rtol=2e-6)


def test_low_rate_regime_bits(gc):
    """Small sigma, y close to mu: almost every likelihood is 1 - tiny and the image costs
    ~1e-5 bit/element.  The fp32 likelihood itself is quantised to 6e-8 there, so the rate of
    ANY fp32 implementation (the reference included) carries percent-level noise; check that
    the fused rate tracks the oracle at that level and the exact value no worse than it does."""
    g = torch.Generator().manual_seed(8)
    shape = (2, 64, 32, 32)
    mu = torch.randn(shape, generator=g)
    sigma = torch.empty(shape).uniform_(0.02, 0.12, generator=g)
    y = mu + 0.05 * torch.randn(shape, generator=g)
    r = ops.gc_forward(y.to(DEV), sigma.to(DEV), mu.to(DEV), want=("lik", "bits"))
    _, lik_ref = cr.gc_forward(y, sigma, mu)
    assert_lik_close(r.lik, lik_ref)
    bits_ref = cr.per_image_bits(lik_ref)
    v32 = ((cr.ste_round(y - mu) + mu) - mu).abs()
    exact = cr.per_image_bits(cr.lower_bound(cr.gc_likelihood(v32.double(), sigma.double(), None), 1e-9))
    err_ours = (r.bits.cpu() - exact).abs() / exact
    err_ref = (bits_ref - exact).abs() / exact
    print("low-rate regime: bits/elem", (exact / y[0].numel()).tolist(), "rel err ours", err_ours.tolist(),
          "reference", err_ref.tolist())
    assert bool((err_ours <= torch.clamp(2 * err_ref, min=0.05)).all())


def test_rate_is_deterministic(gc):
    y, mu, sigma, _ = _rand((8, 64, 48, 32), 77)
    yd, md, sd = y.to(DEV), mu.to(DEV), sigma.to(DEV)
    ref = ops.gc_forward(yd, sd, md, want=("bits",)).bits.clone()
    for _ in range(5):
        assert torch.equal(ops.gc_forward(yd, sd, md, want=("bits",)).bits, ref)


@pytest.mark.parametrize("cfg", [2, 3, 4, 5])
def test_full_size_properties(gc, cfg):
    """BASELINE.json sizes: size-independent properties instead of an elementwise oracle —
    sym/ste round trip, index monotone in sigma and consistent with the table, the fused
    rate equals the sum over the returned likelihoods, a 1-image oracle spot check."""
    c = synthetic.CONFIGS[cfg]
    B = min(c.batch, 64)
    batch = synthetic.make_batch(cfg, range(B), with_noise=c.training)
    y, mu, sg = (batch[k].to(DEV) for k in ("y", "mu", "sigma"))
    table = synthetic.scale_table(DEV)
    gc.scale_table = table
    with torch.no_grad():
        r = gc.forward_fused(y, sg, mu, training=False, want=("ste", "lik", "sym", "idx", "bits"))
    assert torch.equal(r.sym.float() + mu, r.ste)
    s = torch.clamp(sg, min=0.11)
    lo = torch.where(r.idx > 0, table[(r.idx - 1).clamp(min=0).long()], torch.zeros_like(s))
    hi = torch.where(r.idx < 63, table[r.idx.long().clamp(max=62)], torch.full_like(s, float("inf")))
    assert bool(((lo < s) & (s <= hi)).all()), "table[idx-1] < max(sigma,0.11) <= table[idx]"
    own = -(torch.log2(r.lik.double()).reshape(B, -1).sum(1))
    assert torch.allclose(r.bits, own, rtol=2e-6)
    assert float(r.lik.min()) >= 1e-9 and float(r.lik.max()) <= 1.0
    one = {k: batch[k][:1] for k in ("y", "mu", "sigma")}
    yhat_ref, lik_ref = cr.gc_forward(one["y"], one["sigma"], one["mu"])
    assert_equal_exact(r.ste[:1], yhat_ref)
    assert_lik_close(r.lik[:1], lik_ref)
    assert torch.allclose(r.bits[:1].cpu(), cr.per_image_bits(lik_ref), rtol=1e-5)


@pytest.mark.gpu
def test_next_slice_prefetch_hint_changes_no_result():
    """reslic_gc_desc.next_y / reslic_eb_desc.next_y are L2 prefetch hints: any value (the next channel slice, a
    misaligned view, a tensor of another shape) leaves every output bit-identical."""
    import torch
    from reslic_tcm_b200 import ops, synthetic

    dev = "cuda:0"
    g = torch.Generator().manual_seed(12)
    y = torch.randn(6, 128, 16, 8, generator=g).to(dev) * 3
    mu = torch.randn(6, 128, 16, 8, generator=g).to(dev)
    sg = (torch.rand(6, 128, 16, 8, generator=g) * 4 + 0.05).to(dev)
    tab = synthetic.scale_table(dev)
    want = ("ste", "lik", "sym", "idx", "bits")
    base = ops.gc_forward(y[:, :64], sg[:, :64], mu[:, :64], want=want, scale_table=tab)
    for hint in (y[:, 64:], y[:, 64:].reshape(-1)[1:], y[:, :3], None):
        r = ops.gc_forward(y[:, :64], sg[:, :64], mu[:, :64], want=want, scale_table=tab, next_y=hint)
        for name in ("ste", "lik", "sym", "idx", "bits"):
            assert torch.equal(getattr(r, name), getattr(base, name)), name


@pytest.mark.gpu
def test_next_channel_slice_inference():
    """ops.next_channel_slice: the module-API path (GaussianConditional.forward on y.chunk(5, 1)[k]) finds the next
    chunk of the same latent for the L2 prefetch hint, and only ever memory inside the same storage."""
    dev = "cuda:0"
    y = torch.randn(2, 320, 4, 3, device=dev)
    chunks = y.chunk(5, 1)
    for k in range(4):
        nxt = ops.next_channel_slice(chunks[k])
        assert nxt is not None and nxt.data_ptr() == chunks[k + 1].data_ptr() and torch.equal(nxt, chunks[k + 1])
    assert ops.next_channel_slice(chunks[4]) is None                      # last slice: nothing behind it
    assert ops.next_channel_slice(y) is None                              # not a slice
    assert ops.next_channel_slice(y[:, :, :2]) is None                    # spatial slicing: not image-major
    assert ops.next_channel_slice(y[:1].chunk(5, 1)[1]) is not None       # B = 1 keeps the row stride
    assert ops.next_channel_slice(torch.randn(2, 64, 4, 3, device=dev)) is None
    # the module path with the inferred hint gives the results of the explicit-slice path
    gc = GaussianConditional(None).to(dev).eval()
    sg = torch.rand_like(y) * 3 + 0.05
    mu = torch.randn_like(y)
    a = gc(chunks[1], sg.chunk(5, 1)[1], means=mu.chunk(5, 1)[1])
    b = gc(chunks[1].contiguous(), sg.chunk(5, 1)[1].contiguous(), means=mu.chunk(5, 1)[1].contiguous())
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
