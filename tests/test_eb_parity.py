"""GPU parity: fused EntropyBottleneck kernel (through the C ABI) vs the CPU oracle and the
golden vectors produced by the reference's own EntropyBottleneckStanh."""
import pytest
import torch

from oracle import compressai_ref as cr
from reslic_tcm_b200 import EntropyBottleneck, ops, synthetic
from tests.util import assert_equal_exact, assert_lik_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(C, trained_like, seed=1234):
    params = synthetic.eb_parameters(C, trained_like=trained_like, seed=seed)
    mod = EntropyBottleneck(C).to(DEV).eval()
    synthetic.load_eb_parameters(mod, params)
    ref = cr.EntropyBottleneckRef(C)
    ref.matrices = [params[f"_matrix{i}"] for i in range(5)]
    ref.biases = [params[f"_bias{i}"] for i in range(5)]
    ref.factors = [params[f"_factor{i}"] for i in range(4)]
    ref.quantiles = params["quantiles"]
    return mod, ref


def _check(mod, ref, z, training=False, noise=None, what=""):
    with torch.no_grad():
        r = mod.forward_fused(z.to(DEV), training=training, want=("zhat", "ste", "lik", "sym", "bits"),
                              noise=None if noise is None else noise.to(DEV))
    zhat_ref, lik_ref = ref.forward(z, training=training, noise=noise)
    med = ref._get_medians().reshape(1, -1, *([1] * (z.dim() - 2)))
    assert_equal_exact(r.zhat, zhat_ref, what + " z_hat")
    assert_equal_exact(r.ste, cr.ste_round(z - med) + med, what + " ste_round(z - med) + med (tcm.py:431-433)")
    assert_equal_exact(r.sym, ref.symbols(z), what + " symbols")
    # fp64 evaluation of the same parameters: the CUDA kernel must be as close to it as the
    # fp32 reference is (both are fp32 evaluations of a 5-layer MLP; they differ in fma use)
    _, lik64 = ref.to(torch.float64).forward(z.double(), training=training,
                                             noise=None if noise is None else noise.double())
    assert_lik_close(r.lik, lik_ref, what=what + " likelihood vs fp32 oracle")
    assert_lik_close(r.lik, lik64.float(), what=what + " likelihood vs fp64 oracle")
    bits_ref = cr.per_image_bits(lik_ref)
    assert torch.allclose(r.bits.cpu(), bits_ref, rtol=1e-5), (r.bits.cpu(), bits_ref)
    own = -(torch.log2(r.lik.double()).reshape(z.shape[0], -1).sum(1)).cpu()
    assert torch.allclose(r.bits.cpu(), own, rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("trained_like", [False, True])
@pytest.mark.parametrize("shape", [(1, 192, 4, 4), (3, 192, 12, 8), (2, 7, 3, 5), (2, 192, 1, 1), (1, 3, 33, 9),
                                   (5, 1, 2, 2)])
def test_eval_forward(shape, trained_like):
    mod, ref = _pair(shape[1], trained_like)
    g = torch.Generator().manual_seed(sum(shape))
    z = 2.0 * torch.randn(shape, generator=g)
    _check(mod, ref, z, what=f"eval {shape} trained={trained_like}")


@pytest.mark.parametrize("shape", [(1, 192, 4, 4), (2, 64, 6, 10)])
def test_noise_forward_explicit_noise(shape):
    mod, ref = _pair(shape[1], True)
    g = torch.Generator().manual_seed(5 + sum(shape))
    z = 2.0 * torch.randn(shape, generator=g)
    noise = torch.empty(shape).uniform_(-0.5, 0.5, generator=g)
    _check(mod, ref, z, training=True, noise=noise, what=f"noise {shape}")


def test_far_tails_and_ties():
    mod, ref = _pair(4, True)
    z = torch.tensor([-0.0, 0.0, 0.5, -0.5, 1.5, 2.5, 30.0, -30.0, 80.0, -80.0, 1e3, -1e3, 0.49999997, 7.25]
                     ).repeat(1, 4, 1).reshape(1, 4, 14, 1)
    _check(mod, ref, z, what="tails/ties")


def test_golden_vectors_from_reference_module():
    g = load_golden("eb_golden.npz")
    C = g["z"].shape[1]
    mod = EntropyBottleneck(C).to(DEV).eval()
    with torch.no_grad():
        for i in range(5):
            getattr(mod, f"_matrix{i}").copy_(g[f"_matrix{i}"])
            getattr(mod, f"_bias{i}").copy_(g[f"_bias{i}"])
            if i < 4:
                getattr(mod, f"_factor{i}").copy_(g[f"_factor{i}"])
        zhat, lik = mod(g["z"].to(DEV), training=False)
    assert_equal_exact(zhat, g["zhat"], "z_hat vs reference EntropyBottleneckStanh")
    assert_lik_close(lik, g["lik"], what="z likelihood vs reference EntropyBottleneckStanh")


def test_philox_noise_mode_consistency():
    mod, ref = _pair(192, True)
    g = torch.Generator().manual_seed(9)
    z = 2.0 * torch.randn((4, 192, 8, 8), generator=g)
    torch.manual_seed(123)
    with torch.no_grad():
        r = mod.forward_fused(z.to(DEV), training=True, want=("zhat", "lik"))
    u = (r.zhat.cpu() - z)
    assert float(u.abs().max()) <= 0.5 + 1e-6 and abs(float(u.mean())) < 5e-3
    lik_ref = cr.lower_bound(ref._likelihood(r.zhat.cpu().permute(1, 0, 2, 3).reshape(192, 1, -1)), 1e-9)
    lik_ref = lik_ref.reshape(192, 4, 8, 8).permute(1, 0, 2, 3)
    assert_lik_close(r.lik, lik_ref, what="philox-mode z likelihood")


def test_philox_field_is_the_same_in_every_kernel_form():
    """One Philox call per four elements (the quad form of eb_fwd_fast_kernel: each lane of a quad evaluates the counter
    of a different image and a 4 x 4 transpose redistributes the words) must draw exactly the field of the per-element
    form — the mirror-math kernel and the backward kernel regenerate it element by element.  Shapes: CTAs with 6-7
    images each (two word groups, the second one partial), and a spatial size that is no multiple of 4 (no quad form)."""
    from reslic_tcm_b200 import _cabi

    mod, _ = _pair(192, True)
    for shape in ((300, 192, 4, 4), (150, 192, 2, 4), (9, 192, 3, 3)):
        z = 2.0 * torch.randn(shape, generator=torch.Generator().manual_seed(4))
        zd = z.to(DEV)
        m, b, f = mod._params()
        try:
            fast = ops.eb_forward(zd, m, b, f, mod._medians_flat(), training=True, want=("zhat",), seed=77, offset=24)
            _cabi.set_math_mode(_cabi.MATH_MIRROR)
            mirror = ops.eb_forward(zd, m, b, f, mod._medians_flat(), training=True, want=("zhat",), seed=77, offset=24)
        finally:
            _cabi.set_math_mode(_cabi.MATH_FAST)
        assert torch.equal(fast.zhat, mirror.zhat), shape
        u = fast.zhat.cpu() - z
        assert float(u.abs().max()) <= 0.5 + 1e-6 and abs(float(u.mean())) < 5e-3
        other = ops.eb_forward(zd, m, b, f, mod._medians_flat(), training=True, want=("zhat",), seed=77, offset=25)
        assert not torch.equal(other.zhat, fast.zhat)


def test_full_size_config2_properties():
    c = synthetic.CONFIGS[2]
    batch = synthetic.make_batch(2, range(c.batch))
    mod, ref = _pair(192, True)
    z = batch["z"]
    with torch.no_grad():
        r = mod.forward_fused(z.to(DEV), want=("ste", "lik", "sym", "bits"))
    med = mod.quantiles[:, 0, 1].reshape(1, -1, 1, 1)
    assert torch.equal(r.sym.float() + med, r.ste)
    own = -(torch.log2(r.lik.double()).reshape(c.batch, -1).sum(1))
    assert torch.allclose(r.bits, own, rtol=2e-6)
    _, lik_ref = ref.forward(z[:2])
    assert_lik_close(r.lik[:2], lik_ref)


def test_cached_eval_table_is_bit_identical_and_tracks_parameters():
    """reslic_eb_build_lut_f32: the table built once equals the one every launch builds for itself (same
    code path), so likelihoods and symbols with and without `lut` are bit-identical (bits to fp32 summation order); the module
    rebuilds its cached table when a parameter changes in place."""
    mod, ref = _pair(192, True, seed=77)
    z = (torch.randn(3, 192, 12, 8, generator=torch.Generator().manual_seed(5)) * 6.0).to(DEV)
    z[0, 0, 0, 0] = 40.0          # beyond the table: direct evaluation
    m, b, f = mod._params()
    want = ("zhat", "lik", "sym", "bits")
    with torch.no_grad():
        plain = ops.eb_forward(z, m, b, f, mod._medians_flat(), want=want, likelihood_bound=1e-9)
        lut = ops.eb_build_lut(m, b, f, mod._medians_flat(), likelihood_bound=1e-9)
        fast = ops.eb_forward(z, m, b, f, mod._medians_flat(), want=want, likelihood_bound=1e-9, lut=lut)
    assert lut.shape == (192, 130)
    for name in ("zhat", "lik", "sym"):
        assert torch.equal(getattr(plain, name), getattr(fast, name)), name
    # the table launch sums the fp32 partials per (image, 8 channels) instead of per (image, channel)
    assert torch.allclose(plain.bits, fast.bits, rtol=2e-6)
    # the table is the likelihood of the integer offsets about the median
    k = torch.arange(-32, 33, device=DEV, dtype=torch.float32)
    zz = (mod._medians_flat()[None, :, None] + k[None, None, :]).contiguous()          # [1, C, 65]
    with torch.no_grad():
        direct = ops.eb_forward(zz, m, b, f, mod._medians_flat(), want=("lik",), likelihood_bound=1e-9).lik[0]
    assert torch.equal(direct, lut[:, :65])
    assert torch.equal(torch.log2(lut[:, :65]).float(), lut[:, 65:]) or torch.allclose(torch.log2(lut[:, :65]), lut[:, 65:], rtol=2e-7, atol=1e-7)
    # cache invalidation
    t0 = mod._eval_lut()
    assert mod._eval_lut() is t0
    with torch.no_grad():
        mod._bias0.add_(0.25)
    t1 = mod._eval_lut()
    assert t1 is not t0 and not torch.equal(t0, t1)
    with torch.no_grad():
        r = mod.forward_fused(z, training=False, want=("lik",))
        again = ops.eb_forward(z, *mod._params(), mod._medians_flat(), want=("lik",), likelihood_bound=1e-9)
    assert torch.equal(r.lik, again.lik)
    with pytest.raises(ValueError):
        ops.eb_forward(z, m, b, f, mod._medians_flat(), want=("lik",), lut=lut[:, :100].contiguous())


def test_unsupported_filters_fail_loudly():
    mod = EntropyBottleneck(4, filters=(3, 3)).to(DEV)
    with pytest.raises(RuntimeError, match="filters"):
        with torch.no_grad():
            mod(torch.zeros(1, 4, 2, 2, device=DEV))
