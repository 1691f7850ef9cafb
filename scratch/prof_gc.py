"""Profiling driver: a few launches of the fused GC kernel on one config-2 slice."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reslic_tcm_b200 import ops, synthetic
dev = "cuda:0"
cfg = int(os.environ.get("CFG", "2"))
c = synthetic.CONFIGS[cfg]
B = c.batch
g = torch.Generator(device=dev).manual_seed(1)
shape = (B, 320 if os.environ.get("WHOLE") else 64, *c.y_hw)
mu = torch.randn(shape, device=dev, generator=g)
sigma = torch.exp(torch.empty(shape, device=dev).uniform_(-3.0, 4.16, generator=g))
y = mu + sigma * torch.randn(shape, device=dev, generator=g)
table = synthetic.scale_table(dev)
want = ["ste", "lik", "bits"] + (["sym", "idx"] if c.with_indexes else []) + (["yhat"] if c.training else [])
out = {}
for i in range(int(os.environ.get("N", "10"))):
    r = ops.gc_forward(y, sigma, mu, training=c.training, want=want, scale_table=table, seed=1, offset=i)
    out = {k: getattr(r, k) for k in want}
torch.cuda.synchronize()
print("done", float(r.bits.sum()))
