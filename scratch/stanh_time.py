"""Development timing of the STanH kernels, the noise-mode bottleneck and the backward kernels on config-5 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reslic_tcm_b200 import ops, synthetic, EntropyBottleneck
from reslic_tcm_b200.stanh import GaussianConditionalStanh, compute_gap
dev = "cuda:0"
PEAK = 6537.6

def timeit(name, fn, n, bpe, launches=12, reps=10):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(launches): fn(i)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (launches * reps)
    print(f"{name:58s}: {us:8.2f} us  {n / us * 1e-3:7.1f} Gelem/s  {n * bpe / us * 1e-3:6.0f} GB/s  {n * bpe / us * 1e-3 / PEAK * 100:5.1f}%", flush=True)

def stanh_cases(B, C, h, w):
    g = torch.Generator(device=dev).manual_seed(1)
    sets = []
    for _ in range(3):
        mu = torch.randn(B, C, h, w, device=dev, generator=g)
        sigma = torch.exp(torch.empty(B, C, h, w, device=dev).uniform_(-3.0, 4.16, generator=g))
        y = mu + sigma * torch.randn(B, C, h, w, device=dev, generator=g)
        gy = torch.randn(B, C, h, w, device=dev, generator=g); gl = torch.randn(B, C, h, w, device=dev, generator=g)
        sets.append((y, sigma, mu, gy, gl))
    n = B * C * h * w
    for sym in (False, True):
        for beta in (10.0, 1.0, 3.0, -1.0):
            cfg = {"beta": beta, "num_sigmoids": 0, "extrema": 80, "symmetry": sym, "trainable": False, "removing_mean": True}
            m = GaussianConditionalStanh(None, channels=C, gaussian_configuration=cfg).to(dev)
            m.stanh.update_state(torch.device(dev))
            def fwd(i, training=True):
                y, s, mu, _, _ = sets[i % 3]
                return m.forward_fused(y, s, training=training, means=mu, want=("yhat", "lik"))
            timeit(f"stanh fwd train sym={int(sym)} beta={beta} [{B},{C},{h},{w}]", fwd, n, 20)
            if beta == 10.0:
                timeit(f"stanh fwd eval(hard) sym={int(sym)} [{B},{C},{h},{w}]", lambda i: fwd(i, False), n, 20)
                timeit(f"stanh symbols sym={int(sym)}", lambda i: m._stanh_fused(sets[i % 3][0], None, sets[i % 3][2], False, ("sym",)), n, 12)
            def bwd(i):
                y, s, mu, gy, gl = sets[i % 3]
                return m._stanh_backward(y, s, mu, True, gy, gl)
            timeit(f"stanh bwd train sym={int(sym)} beta={beta}", bwd, n, 32)
            if beta in (10.0, 1.0):
                timeit(f"compute_gap sym={int(sym)} beta={beta}", lambda i: compute_gap(m.stanh, sets[i % 3][0]), n, 4)

def eb_cases(B, C, h, w):
    mod = EntropyBottleneck(C).to(dev).train()
    synthetic.load_eb_parameters(mod, synthetic.eb_parameters())
    m, b, f = mod._params(); med = mod._medians_flat()
    zs = [torch.randn(B, C, h, w, device=dev) * 4 for _ in range(3)]
    gs = [torch.randn(B, C, h, w, device=dev) for _ in range(3)]
    n = B * C * h * w
    timeit(f"eb fwd noise [{B},{C},{h},{w}]", lambda i: ops.eb_forward(zs[i % 3], m, b, f, med, training=True, want=("zhat", "lik"), seed=1, offset=i), n, 12)
    timeit(f"eb fwd eval direct [{B},{C},{h},{w}]", lambda i: ops.eb_forward(zs[i % 3], m, b, f, med, training=False, want=("zhat", "lik")), n, 12)
    try:
        timeit(f"eb bwd noise [{B},{C},{h},{w}]", lambda i: ops.eb_backward(zs[i % 3], m, b, f, med, training=True, g_zhat=gs[i % 3], g_lik=gs[(i + 1) % 3], seed=1, offset=i), n, 20)
    except Exception as e:
        print("eb bwd:", repr(e)[:300])

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "stanh"):
    stanh_cases(256, 64, 16, 16)
if which in ("all", "eb"):
    eb_cases(256, 192, 4, 4)
    eb_cases(24, 192, 12, 8)
if which == "once":
    # one eager call of each op (for an ncu launch list)
    B, C, h, w = 256, 64, 16, 16
    g = torch.Generator(device=dev).manual_seed(1)
    mu = torch.randn(B, C, h, w, device=dev, generator=g)
    sigma = torch.exp(torch.empty(B, C, h, w, device=dev).uniform_(-3.0, 4.16, generator=g))
    y = mu + sigma * torch.randn(B, C, h, w, device=dev, generator=g)
    cfg = {"beta": 10.0, "num_sigmoids": 0, "extrema": 80, "symmetry": False, "trainable": False, "removing_mean": True}
    m = GaussianConditionalStanh(None, channels=C, gaussian_configuration=cfg).to(dev)
    m.stanh.update_state(torch.device(dev))
    for _ in range(3):
        m.forward_fused(y, sigma, training=True, means=mu, want=("yhat", "lik"))
        m.forward_fused(y, sigma, training=False, means=mu, want=("yhat", "lik"))
        m._stanh_backward(y, sigma, mu, True, y, sigma)
        compute_gap(m.stanh, y)
    torch.cuda.synchronize()
if which == "once_eb":
    B, C, h, w = 256, 192, 4, 4
    mod = EntropyBottleneck(C).to(dev).train()
    synthetic.load_eb_parameters(mod, synthetic.eb_parameters())
    m, b, f = mod._params(); med = mod._medians_flat()
    z = torch.randn(B, C, h, w, device=dev) * 4
    gz = torch.randn(B, C, h, w, device=dev)
    for i in range(3):
        ops.eb_forward(z, m, b, f, med, training=True, want=("zhat", "lik"), seed=1, offset=i)
        ops.eb_backward(z, m, b, f, med, training=True, g_zhat=gz, g_lik=gz, seed=1, offset=i)
    torch.cuda.synchronize()
