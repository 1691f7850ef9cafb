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
