"""Development timing of the Gaussian-conditional backward (and forward) on config-5-shaped slices."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reslic_tcm_b200 import ops
dev = "cuda:0"
def run(B, C, h, w, training):
    g = torch.Generator(device=dev).manual_seed(1)
    sets = []
    for _ in range(3):
        mu = torch.randn(B, C, h, w, device=dev, generator=g)
        sigma = torch.exp(torch.empty(B, C, h, w, device=dev).uniform_(-3.0, 4.16, generator=g))
        y = mu + sigma * torch.randn(B, C, h, w, device=dev, generator=g)
        gy = torch.randn(B, C, h, w, device=dev, generator=g); gl = torch.randn(B, C, h, w, device=dev, generator=g)
        sets.append((y, sigma, mu, gy, gl))
    def one(i):
        y, s, m, gy, gl = sets[i % 3]
        return ops.gc_backward(y, s, m, training=training, g_yhat=gy if training else None, g_ste=None if training else gy,
                               g_lik=gl, seed=3, offset=i % 3)
    for i in range(3): one(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(15): one(i)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 150
    n = B * C * h * w
    bpe = 12 + 8 + 12     # y, mu, sigma + g_yhat|g_ste, g_lik -> g_y, g_mu, g_sigma
    print(f"bwd B={B} C={C} {h}x{w} training={training}: {us:.2f} us  {n * bpe / us * 1e-3:.0f} GB/s  {n * bpe / us * 1e-3 / 65.376:.1f}% of 6537.6")
run(256, 64, 16, 16, True)
run(256, 320, 16, 16, True)
run(24, 64, 48, 32, False)
run(24, 320, 48, 32, False)
