"""Development timing of the bottleneck launch alone (graph of back-to-back launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reslic_tcm_b200 import ops, synthetic, _cabi, EntropyBottleneck
dev = "cuda:0"
mod = EntropyBottleneck(192).to(dev).eval()
synthetic.load_eb_parameters(mod, synthetic.eb_parameters())
m, b, f = mod._params()
med = mod._medians_flat()
B = 24
zs = [torch.randn(B, 192, 12, 8, device=dev) * 4 for _ in range(3)]
outs = [{"ste": torch.empty_like(zs[0]), "lik": torch.empty_like(zs[0])} for _ in range(3)]
ws = torch.zeros(int(_cabi.load().reslic_workspace_bytes(B)), dtype=torch.uint8, device=dev)
bits = torch.zeros(B, dtype=torch.float64, device=dev)
lut = mod._eval_lut()

def run(mode, use_lut, want=("ste", "lik", "bits")):
    def one(i):
        out = dict(outs[i % 3], workspace=ws)
        if mode == "deferred":
            out["bits_deferred"] = True
        elif mode == "collect":
            out["bits"], out["bits_collect"] = bits, True
        else:
            out["bits"] = bits
        w = want if mode != "none" else tuple(x for x in want if x != "bits")
        ops.eb_forward(zs[i % 3], m, b, f, med, want=w, out=out, lut=lut if use_lut else None)
    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(30):
            one(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"mode={mode:9s} lut={int(use_lut)} want={','.join(want):14s}: {e0.elapsed_time(e1) * 1e3 / 600:.2f} us/launch")

for mode in ("immediate", "deferred", "collect", "none"):
    for use_lut in (False, True):
        run(mode, use_lut)
run("none", True, want=("lik",))
