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


def test_struct_sizes_match_header_layout():
    # 64-bit ABI: computed by hand from the header's field order
    assert ctypes.sizeof(_cabi.GcDesc) % 8 == 0
    assert _cabi.GcDesc.B.offset == 64 and _cabi.GcDesc.mode.offset == 80
    assert _cabi.EbDesc.matrix.offset == 64 and _cabi.EbDesc.medians.offset == 64 + 14 * 8


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
