"""GPU: out-of-bounds write detection with guard bands (compute-sanitizer is not available on this
pool): every output buffer is an interior view of a larger canary-filled allocation."""
import pytest
import torch

from reslic_tcm_b200 import ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CANARY = -1234.5


def _guarded(shape, dtype, pad=64):
    n = 1
    for s in shape:
        n *= s
    buf = torch.full((n + 2 * pad,), CANARY if dtype.is_floating_point else -77, dtype=dtype, device=DEV)
    return buf, buf[pad:pad + n].view(shape), pad, n


def _intact(buf, pad, n):
    fill = CANARY if buf.dtype.is_floating_point else -77
    return bool((buf[:pad] == fill).all()) and bool((buf[pad + n:] == fill).all())


@pytest.mark.parametrize("shape", [(3, 64, 8, 8), (2, 5, 3, 7), (1, 1, 1, 1), (4, 3, 1, 5), (2, 64, 48, 32)])
@pytest.mark.parametrize("training", [False, True])
def test_gc_outputs_stay_inside_their_buffers(shape, training):
    g = torch.Generator().manual_seed(sum(shape))
    y = torch.randn(shape, generator=g).to(DEV)
    mu = torch.randn(shape, generator=g).to(DEV)
    sg = (torch.rand(shape, generator=g) * 3 + 0.05).to(DEV)
    outs, bufs = {}, {}
    for name, dt in (("yhat", torch.float32), ("ste", torch.float32), ("lik", torch.float32), ("sym", torch.int32),
                     ("idx", torch.int32)):
        bufs[name] = _guarded(shape, dt)
        outs[name] = bufs[name][1]
    bbuf, bits, bpad, bn = _guarded((shape[0],), torch.float64)
    outs["bits"] = bits
    ops.gc_forward(y, sg, mu, training=training, want=("yhat", "ste", "lik", "sym", "idx", "bits"), out=outs,
                   scale_table=synthetic.scale_table(DEV), seed=3)
    torch.cuda.synchronize()
    for name, (buf, view, pad, n) in bufs.items():
        assert _intact(buf, pad, n), f"{name}: guard band overwritten"
        assert bool((view != (CANARY if view.dtype.is_floating_point else -77)).any()), f"{name}: nothing written"
    assert _intact(bbuf, bpad, bn)
    # backward outputs
    gy, gm, gs = ops.gc_backward(y, sg, mu, training=training, g_yhat=torch.ones_like(y), g_lik=torch.ones_like(y),
                                 g_ste=None, seed=3)
    assert gy.shape == y.shape and torch.isfinite(gy).all() and torch.isfinite(gs).all()


@pytest.mark.parametrize("shape", [(2, 192, 4, 4), (3, 7, 3, 5), (1, 3, 33, 9), (2, 192, 1, 1)])
@pytest.mark.parametrize("training", [False, True])
def test_eb_outputs_stay_inside_their_buffers(shape, training):
    C = shape[1]
    params = synthetic.eb_parameters(C, trained_like=True)
    dev = {k: v.to(DEV) for k, v in params.items()}
    g = torch.Generator().manual_seed(sum(shape) + 1)
    z = (2.0 * torch.randn(shape, generator=g)).to(DEV)
    outs, bufs = {}, {}
    for name, dt in (("zhat", torch.float32), ("ste", torch.float32), ("lik", torch.float32), ("sym", torch.int32)):
        bufs[name] = _guarded(shape, dt)
        outs[name] = bufs[name][1]
    ops.eb_forward(z, [dev[f"_matrix{i}"] for i in range(5)], [dev[f"_bias{i}"] for i in range(5)],
                   [dev[f"_factor{i}"] for i in range(4)], dev["quantiles"][:, 0, 1].contiguous(), training=training,
                   want=("zhat", "ste", "lik", "sym", "bits"), out=outs, seed=5)
    torch.cuda.synchronize()
    for name, (buf, view, pad, n) in bufs.items():
        assert _intact(buf, pad, n), f"{name}: guard band overwritten"


@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (3, 5, 7, 3), (1, 1, 1, 1), (2, 3, 4, 4), (5, 8, 1, 4)])
@pytest.mark.parametrize("beta,training", [(10.0, True), (-1, True), (10.0, False)])
def test_stanh_outputs_stay_inside_their_buffers(shape, beta, training):
    """The 128-bit STanH forward (aligned shapes) and the scalar one (odd shapes): outputs in guarded buffers."""
    from reslic_tcm_b200.stanh import GaussianConditionalStanh

    cfg = dict(beta=beta, num_sigmoids=0, extrema=20, trainable=False, removing_mean=True, symmetry=False)
    m = GaussianConditionalStanh(None, channels=shape[1], gaussian_configuration=cfg).to(DEV)
    m.stanh.update_state(torch.device(DEV))
    g = torch.Generator().manual_seed(sum(shape) + 2)
    y = (8.0 * torch.randn(shape, generator=g)).to(DEV)
    mu = torch.randn(shape, generator=g).to(DEV)
    sg = (torch.rand(shape, generator=g) * 3 + 0.05).to(DEV)
    outs, bufs = {}, {}
    for name, dt in (("yhat", torch.float32), ("lik", torch.float32), ("sym", torch.int32)):
        bufs[name] = _guarded(shape, dt)
        outs[name] = bufs[name][1]
    bbuf, bits, bpad, bn = _guarded((shape[0],), torch.float64)
    outs["bits"] = bits
    m.forward_fused(y, sg, training=training, means=mu, want=("yhat", "lik", "sym", "bits"), out=outs)
    torch.cuda.synchronize()
    for name, (buf, view, pad, n) in bufs.items():
        assert _intact(buf, pad, n), f"{name}: guard band overwritten"
        assert bool((view != (CANARY if view.dtype.is_floating_point else -77)).any()), f"{name}: nothing written"
    assert _intact(bbuf, bpad, bn) and torch.isfinite(bits).all()
    # same values as the kernel's own allocation path
    r = m.forward_fused(y, sg, training=training, means=mu, want=("yhat", "lik", "sym"))
    assert torch.equal(r["yhat"], outs["yhat"]) and torch.equal(r["sym"], outs["sym"])


@pytest.mark.parametrize("n", [1, 5, 4096, 4099])
def test_rans_slots_and_rate_outputs_stay_inside_their_buffers(n):
    from oracle import compressai_ref as cr

    cdf, offset, length = cr.gc_update(cr.get_scale_table())
    g = torch.Generator().manual_seed(n)
    sym = torch.randint(-40, 40, (n,), generator=g, dtype=torch.int32).to(DEV)
    idx = torch.randint(0, 64, (n,), generator=g, dtype=torch.int32).to(DEV)
    sbuf, slots, spad, sn = _guarded((n,), torch.int32)
    cap = 16
    pbuf, epos, ppad, pn = _guarded((cap,), torch.int32)
    rbuf, eraw, rpad, rn = _guarded((cap,), torch.int64)
    tbuf, status, tpad, tn = _guarded((2,), torch.int32)
    ops.rans_slots(sym, idx, cdf.to(DEV), length.to(DEV), offset.to(DEV), out=(slots, epos, eraw, status))
    torch.cuda.synchronize()
    for buf, pad, nn_ in ((sbuf, spad, sn), (pbuf, ppad, pn), (rbuf, rpad, rn), (tbuf, tpad, tn)):
        assert _intact(buf, pad, nn_)
    assert int(status[1]) == 0
    lik = (torch.rand(3, n, generator=g) * 0.9 + 1e-6).to(DEV)
    bbuf, bits, bpad, bn = _guarded((3,), torch.float64)
    ops.rate_from_likelihood(lik, out={"bits": bits})
    torch.cuda.synchronize()
    assert _intact(bbuf, bpad, bn)
    assert torch.allclose(bits, -(torch.log2(lik.double()).sum(1)), rtol=2e-6, atol=2e-5)    # 48.16 fixed point: 2^-16 bit
