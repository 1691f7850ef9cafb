"""dev: noise-mode bottleneck forward/backward launch time vs RESLIC_EB_SPLITS (config-5 z: 256 x 192 x 4 x 4)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from reslic_tcm_b200 import EntropyBottleneck, ops, synthetic

dev = torch.device("cuda:0")
B, Cz, hz = 256, 192, 4
g = torch.Generator(device=dev).manual_seed(1)
mod = EntropyBottleneck(Cz).to(dev).train()
synthetic.load_eb_parameters(mod, synthetic.eb_parameters())
mm, bb, ff = mod._params()
med = mod._medians_flat()
zs = [torch.randn(B, Cz, hz, hz, device=dev, generator=g) * 4 for _ in range(3)]
bits = torch.zeros(B, device=dev, dtype=torch.float64)


def timeit(fn, launches=12, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.graph(gr):
        for i in range(launches):
            keep.append(fn(i))
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (launches * reps)


for sp in sys.argv[1:] or ["0"]:
    if sp == "0":
        os.environ.pop("RESLIC_EB_SPLITS", None)
    else:
        os.environ["RESLIC_EB_SPLITS"] = sp
    a = timeit(lambda i: ops.eb_forward(zs[i % 3], mm, bb, ff, med, training=True, want=("zhat", "lik"), seed=1, offset=i))
    b = timeit(lambda i: ops.eb_forward(zs[i % 3], mm, bb, ff, med, training=True, want=("zhat", "lik", "bits"), seed=1, offset=i))
    print(f"splits={sp:>4}  zhat+lik {a:6.2f} us   +bits {b:6.2f} us", flush=True)
