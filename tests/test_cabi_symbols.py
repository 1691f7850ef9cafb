"""CPU: the C-ABI library builds, loads, and exports every symbol include/reslic_b200.h
declares (no compute calls without a GPU), and the host-side pieces behave."""
import ctypes
import os
import re

import pytest
import torch

from oracle import compressai_ref as cr
from reslic_tcm_b200 import _build, _cabi, cdf_tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _cabi.load()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "reslic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(reslic_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared_functions()
    assert len(names) >= 8
    raw = ctypes.CDLL(_cabi.lib_path())
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
        assert n in _cabi.EXPORTS, f"{n} has no ctypes prototype"
    assert sorted(_cabi.EXPORTS) == names


def test_abi_version_and_workspace(lib):
    assert lib.reslic_abi_version() == _cabi.ABI_VERSION
    assert lib.reslic_workspace_bytes(0) == 0
    assert lib.reslic_workspace_bytes(24) >= 24 * 8
    assert lib.reslic_workspace_bytes(-3) == 0


def test_struct_sizes_match_header_layout(lib):
    # 64-bit ABI: computed by hand from the header's field order (struct_size is the first member of every descriptor)
    assert ctypes.sizeof(_cabi.GcDesc) % 8 == 0
    assert _cabi.GcDesc.struct_size.offset == 0 and _cabi.GcDesc.B.offset == 72 and _cabi.GcDesc.mode.offset == 88
    assert _cabi.EbDesc.matrix.offset == 72 and _cabi.EbDesc.medians.offset == 72 + 14 * 8
    # ... and every ctypes struct is exactly as large as the library's own sizeof
    for cls, fn in _cabi.STRUCT_SIZES.items():
        assert ctypes.sizeof(cls) == getattr(lib, fn)(), cls.__name__
    for cls in _cabi.STRUCT_SIZES:
        if cls is not _cabi.StanhTables:
            assert cls._fields_[0][0] == "struct_size" and _cabi.new(cls).struct_size == ctypes.sizeof(cls)


def test_short_or_stale_descriptor_is_rejected_before_any_field_is_read(lib):
    """A binding built against another ABI revision (the pre-next_y struct of INTEGRATION.md) must get
    RESLIC_ERR_ARG, not a library that reads past its struct.  No GPU needed: the check precedes everything."""
    calls = [(_cabi.GcDesc, lib.reslic_gc_fwd_f32), (_cabi.GcBwdDesc, lib.reslic_gc_bwd_f32),
             (_cabi.EbDesc, lib.reslic_eb_fwd_f32), (_cabi.EbBwdDesc, lib.reslic_eb_bwd_f32),
             (_cabi.StanhGcDesc, lib.reslic_stanh_gc_fwd_f32), (_cabi.StanhGcBwdDesc, lib.reslic_stanh_gc_bwd_f32),
             (_cabi.EbStanhDesc, lib.reslic_eb_stanh_fwd_f32)]
    for cls, fn in calls:
        for bad in (0, ctypes.sizeof(cls) - 16, ctypes.sizeof(cls) + 8):
            d = cls()
            d.struct_size = bad
            assert fn(ctypes.byref(d), None) == -1, (cls.__name__, bad)
            assert b"struct_size" in lib.reslic_last_error()
    d = _cabi.new(_cabi.GcDesc)
    assert lib.reslic_gc_fwd_f32(ctypes.byref(d), None) == 0      # correctly sized, B = 0: empty input, nothing to do
    d = _cabi.new(_cabi.EbDesc)
    d.struct_size -= 8
    assert lib.reslic_eb_build_lut_f32(ctypes.byref(d), ctypes.c_void_p(8), None) == -1


def test_rate_exchange_layout(lib):
    assert lib.reslic_rate_exchange_bytes(8, 256) == 256 * 8 * 64
    assert lib.reslic_rate_exchange_bytes(0, 4) == 0 and lib.reslic_rate_exchange_bytes(2, 0) == 0
    assert lib.reslic_workspace_bytes(24) == (4 * 24 + 4) * 8


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setenv("RESLIC_B200_LIB", "/nonexistent/libreslic_b200.so")
    with pytest.raises(_cabi.ReslicError, match="no fallback"):
        _cabi.load()


def test_cpu_tensors_are_rejected_not_silently_computed(lib):
    from reslic_tcm_b200 import GaussianConditional

    gc = GaussianConditional(None)
    x = torch.zeros(1, 4, 2, 2)
    with pytest.raises(_cabi.ReslicError, match="no CPU fallback"):
        gc(x, x + 1.0, x)


def test_pmf_to_quantized_cdf_matches_oracle(lib):
    g = torch.Generator().manual_seed(11)
    for n in (1, 2, 5, 33, 257, 1200):
        p = torch.rand(n, generator=g) ** 6
        p = p / p.sum()
        got = cdf_tables.pmf_to_quantized_cdf(p, 16).tolist()
        assert got == cr.pmf_to_quantized_cdf(p.tolist(), 16)
    with pytest.raises(ValueError):
        cdf_tables.pmf_to_quantized_cdf(torch.tensor([0.3, -0.2]))
    with pytest.raises(ValueError):
        cdf_tables.pmf_to_quantized_cdf(torch.tensor([0.0, 0.0]))


def test_gaussian_conditional_update_tables_match_oracle(lib):
    from reslic_tcm_b200 import GaussianConditional

    gc = GaussianConditional(None)
    assert gc.update_scale_table(cr.get_scale_table()) is True
    assert gc.update_scale_table(cr.get_scale_table()) is False      # already initialised
    cdf, off, ln = cr.gc_update(cr.get_scale_table())
    assert torch.equal(gc.quantized_cdf, cdf) and torch.equal(gc.offset, off) and torch.equal(gc.cdf_length, ln)


def test_entropy_bottleneck_state_dict_names_and_update(lib):
    from reslic_tcm_b200 import EntropyBottleneck

    torch.manual_seed(0)
    eb = EntropyBottleneck(6)
    names = set(eb.state_dict().keys())
    want = {f"_matrix{i}" for i in range(5)} | {f"_bias{i}" for i in range(5)} | {f"_factor{i}" for i in range(4)}
    want |= {"quantiles", "target", "_offset", "_quantized_cdf", "_cdf_length"}
    assert want <= names
    assert eb._matrix0.shape == (6, 3, 1) and eb._matrix4.shape == (6, 1, 3) and eb.quantiles.shape == (6, 1, 3)
    ref = cr.EntropyBottleneckRef(6)
    ref.matrices = [getattr(eb, f"_matrix{i}").detach() for i in range(5)]
    ref.biases = [getattr(eb, f"_bias{i}").detach() for i in range(5)]
    ref.factors = [getattr(eb, f"_factor{i}").detach() for i in range(4)]
    assert eb.update() is True and eb.update() is False
    cdf, off, ln = ref.update()
    assert torch.equal(eb.quantized_cdf, cdf) and torch.equal(eb.offset, off) and torch.equal(eb.cdf_length, ln)
    assert float(eb.loss().detach()) == pytest.approx(float(ref.loss()), rel=1e-6)


def test_next_channel_slice_is_pure_view_logic():
    """ops.next_channel_slice (the inferred L2 prefetch hint of the module path) on CPU tensors: it finds
    y.chunk(5, 1)[k + 1] behind y.chunk(5, 1)[k] and never describes memory outside the storage."""
    from reslic_tcm_b200 import ops

    y = torch.randn(3, 320, 4, 3)
    ch = y.chunk(5, 1)
    for k in range(4):
        nxt = ops.next_channel_slice(ch[k])
        assert nxt is not None and nxt.data_ptr() == ch[k + 1].data_ptr() and torch.equal(nxt, ch[k + 1])
    assert ops.next_channel_slice(ch[4]) is None and ops.next_channel_slice(y) is None
    assert ops.next_channel_slice(y[:, :64, :2]) is None
    assert ops.next_channel_slice(y[:, 10:74]) is not None and ops.next_channel_slice(y[:, 200:264]) is None   # 264 + 64 > 320
    # a slice of a tensor that sits at the END of a larger storage: the bound is the storage, not the shape
    flat = torch.randn(2 * 128 * 6 + 5)
    t = flat[5:].view(2, 128, 6)
    assert ops.next_channel_slice(t[:, :64]) is None or ops.next_channel_slice(t[:, :64]).data_ptr() == t[:, 64:].data_ptr()
    assert ops.next_channel_slice(y.double()[:, :64]) is None


def test_integration_doc_stub_is_generated_from_the_binding():
    """INTEGRATION.md §3 shows the ctypes struct a maintainer would write; it is generated from _cabi.GcDesc
    (tools/gen_integration_stub.py) so that it cannot lag an ABI revision behind again."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_integration_stub", os.path.join(root, "tools", "gen_integration_stub.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    assert mod.render(text) == text, "INTEGRATION.md is stale: run python tools/gen_integration_stub.py"
    # and the stub as printed really is a correctly sized descriptor for this library
    code = text[text.index(mod.BEGIN):text.index(mod.END)].split("```python\n", 1)[1].rsplit("```", 1)[0]
    code = code.replace('C.CDLL("reslic_tcm_b200/lib/libreslic_b200.so")', "C.CDLL(LIB)")
    ns = {"LIB": _cabi.lib_path()}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    assert ctypes.sizeof(ns["GcDesc"]) == ctypes.sizeof(_cabi.GcDesc)
