"""Development probe: pinned H2D / D2H bandwidth alone and together (sizes of the config-2 e2e step)."""
import torch
dev = "cuda:0"
h_in = torch.empty(143327232 // 4, dtype=torch.float32).pin_memory()
d_in = torch.empty_like(h_in, device=dev)
d_out = torch.empty(94372032 // 4, dtype=torch.int32, device=dev)
h_out = torch.empty(94372032 // 4, dtype=torch.int32).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=10):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
for f in (h2d, d2h):
    f()
ms = t(h2d); print(f"H2D 143 MB alone : {ms:.3f} ms  {143.3 / ms:.1f} GB/s")
ms = t(d2h); print(f"D2H  94 MB alone : {ms:.3f} ms  {94.4 / ms:.1f} GB/s")
ms = t(lambda: (h2d(), d2h())); print(f"both concurrently: {ms:.3f} ms per pair")
