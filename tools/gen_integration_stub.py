#!/usr/bin/env python
"""Regenerates the ctypes stub shown in INTEGRATION.md §3 from reslic_tcm_b200/_cabi.py, so that the document cannot
drift from the binding again (round 1 shipped a stub one ABI revision old).  `--check` exits 1 if the file is stale;
tests/test_cabi_symbols.py runs the same comparison."""
from __future__ import annotations

import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BEGIN, END = "<!-- BEGIN GENERATED gc_desc_stub (tools/gen_integration_stub.py) -->", "<!-- END GENERATED gc_desc_stub -->"

NAMES = {C.c_void_p: "C.c_void_p", C.c_int64: "C.c_int64", C.c_int32: "C.c_int32", C.c_uint64: "C.c_uint64",
         C.c_float: "C.c_float", C.c_double: "C.c_double"}


def _fields(cls) -> str:
    from reslic_tcm_b200 import _cabi

    items = []
    for name, typ in cls._fields_:
        if typ in NAMES:
            t = NAMES[typ]
        elif typ is C.POINTER(_cabi.RateExchangeDesc):
            t = "C.c_void_p"            # const reslic_rate_exchange* (NULL on one GPU)
        else:
            raise SystemExit(f"no spelling for the type of {cls.__name__}.{name}")
        items.append(f'("{name}", {t})')
    lines, cur = [], "    _fields_ = ["
    for it in items:
        if len(cur) + len(it) + 2 > 116:
            lines.append(cur.rstrip())
            cur = "                "
        cur += it + ", "
    lines.append(cur.rstrip(", ") + "]")
    return "\n".join(lines)


def stub() -> str:
    from reslic_tcm_b200 import _cabi

    return f'''```python
import ctypes as C, torch
lib = C.CDLL("reslic_tcm_b200/lib/libreslic_b200.so")
assert lib.reslic_abi_version() == {_cabi.ABI_VERSION}

class GcDesc(C.Structure):            # struct reslic_gc_desc, field for field (ABI {_cabi.ABI_VERSION})
{_fields(_cabi.GcDesc)}

lib.reslic_sizeof_gc_desc.restype = C.c_int64
assert C.sizeof(GcDesc) == lib.reslic_sizeof_gc_desc()      # a stale stub fails here, not inside a kernel
lib.reslic_gc_fwd_f32.restype = C.c_int
lib.reslic_gc_fwd_f32.argtypes = [C.POINTER(GcDesc), C.c_void_p]

def gaussian_conditional_forward(y, scales, means):        # y may be y.chunk(5, 1)[k]: no copy
    d = GcDesc()
    d.struct_size = C.sizeof(GcDesc)                       # checked by the library before any field is read
    d.y, d.y_bs = y.data_ptr(), y.stride(0)
    d.mu, d.mu_bs = means.data_ptr(), means.stride(0)
    d.sigma, d.sigma_bs = scales.data_ptr(), scales.stride(0)
    d.B, d.n = y.shape[0], y[0].numel()
    d.mode, d.scale_bound, d.likelihood_bound = 0, 0.11, 1e-9   # RESLIC_Q_DEQUANTIZE
    y_hat, lik = torch.empty_like(y), torch.empty_like(y)
    d.yhat, d.yhat_bs, d.lik, d.lik_bs = y_hat.data_ptr(), y_hat.stride(0), lik.data_ptr(), lik.stride(0)
    rc = lib.reslic_gc_fwd_f32(C.byref(d), torch.cuda.current_stream().cuda_stream)
    if rc: raise RuntimeError(lib.reslic_last_error().decode())
    return y_hat, lik
```'''


def render(text: str) -> str:
    a, b = text.index(BEGIN), text.index(END)
    return text[:a] + BEGIN + "\n" + stub() + "\n" + text[b:]


if __name__ == "__main__":
    path = os.path.join(ROOT, "INTEGRATION.md")
    old = open(path).read()
    new = render(old)
    if "--check" in sys.argv:
        sys.exit(0 if new == old else 1)
    open(path, "w").write(new)
    print("INTEGRATION.md section 3 regenerated" if new != old else "INTEGRATION.md up to date")
