"""dev: noise-mode bottleneck backward launch time (config-5 z: 256 x 192 x 4 x 4); RESLIC_B200_LIB / RESLIC_EBB_SPLITS from the env."""
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
gs = [torch.randn(B, Cz, hz, hz, device=dev, generator=g) for _ in range(3)]
fn = lambda i: ops.eb_backward(zs[i % 3], mm, bb, ff, med, training=True, g_zhat=gs[i % 3], g_lik=gs[(i + 1) % 3], seed=1, offset=i)
for i in range(3):
    r = fn(i)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
keep = []
with torch.cuda.graph(gr):
    for i in range(12):
        keep.append(fn(i))
gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    gr.replay()
e1.record()
torch.cuda.synchronize()
chk = [float(t.double().abs().sum()) for t in (r if isinstance(r, (tuple, list)) else [r]) if torch.is_tensor(t)]
print(f"lib={os.path.basename(os.environ.get('RESLIC_B200_LIB', 'default'))} splits={os.environ.get('RESLIC_EBB_SPLITS', '-')} "
      f"{e0.elapsed_time(e1) * 1e3 / 240:7.2f} us  chk={chk[:3]}", flush=True)
