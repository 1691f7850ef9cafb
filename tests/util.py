"""Shared comparison helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Likelihood tolerance (north star: 1e-5 relative).  The absolute term covers the
# cancellation regime of the reference formula itself (L = upper - lower with both terms
# O(1) for large sigma; erfc values in [1,2) have a 1.2e-7 ulp): there the reference's OWN
# fp32 result deviates from the exact value by up to ~1.1e-7 absolute (BASELINE.md §4:
# 7.2e-5 relative at sigma=256, L~1.5e-3), so two correct fp32 implementations can differ
# by the sum of their errors.  SURVEY.md §7.4 H2 proposes 2e-7; against the exact (fp64)
# value we hold that (LIK_ATOL_EXACT), against the fp32 oracle we allow 3e-7.
LIK_RTOL = 1e-5
LIK_ATOL = 3e-7
LIK_ATOL_EXACT = 2e-7


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def lik_close(actual: torch.Tensor, ref: torch.Tensor, rtol=LIK_RTOL, atol=LIK_ATOL):
    """|a - r| <= rtol*|r| + atol elementwise, NaN == NaN; returns (ok, worst-violation message)."""
    a, r = actual.double().cpu(), ref.double().cpu()
    nan_ok = torch.isnan(a) == torch.isnan(r)
    a0, r0 = torch.nan_to_num(a), torch.nan_to_num(r)
    err = (a0 - r0).abs()
    tol = rtol * r0.abs() + atol
    bad = (err > tol) | ~nan_ok
    if bad.any():
        i = int(torch.argmax((err - tol) * bad))
        return False, f"{int(bad.sum())} bad; worst at flat {i}: got {a.reshape(-1)[i].item():.9g} ref {r.reshape(-1)[i].item():.9g}"
    return True, f"max abs err {err.max().item():.3g}, max rel err {(err / r0.abs().clamp_min(1e-30)).max().item():.3g}"


def assert_lik_close(actual, ref, rtol=LIK_RTOL, atol=LIK_ATOL, what="likelihood"):
    ok, msg = lik_close(actual, ref, rtol, atol)
    assert ok, f"{what}: {msg}"


def assert_equal_exact(actual: torch.Tensor, ref: torch.Tensor, what="tensor"):
    a, r = actual.cpu(), ref.cpu()
    assert a.shape == r.shape, f"{what}: shape {tuple(a.shape)} != {tuple(r.shape)}"
    assert a.dtype == r.dtype, f"{what}: dtype {a.dtype} != {r.dtype}"
    if a.is_floating_point():
        same = (a == r) | (torch.isnan(a) & torch.isnan(r))
    else:
        same = a == r
    n = int((~same).sum())
    if n:
        i = int(torch.nonzero(~same.reshape(-1))[0])
        raise AssertionError(f"{what}: {n} mismatches; first at flat {i}: got {a.reshape(-1)[i].item()} ref {r.reshape(-1)[i].item()}")


def rans_slots_reference(sym, idx, cdf, length, offset):
    """Host restatement of reslic_rans_slots_u32 (the lookup at the head of compressai.ans.encode_with_indexes,
    SURVEY.md App. A.6): (slots int32, escape positions ascending, escape raw values)."""
    s, i = sym.reshape(-1).long(), idx.reshape(-1).long()
    max_value = (length.reshape(-1).long() - 2)[i]
    value = s - offset.reshape(-1).long()[i]
    neg, big = value < 0, value >= max_value
    raw = torch.where(neg, -2 * value - 1, 2 * (value - max_value))
    value = torch.where(neg | big, max_value, value)
    c = cdf.long()
    start, nxt = c[i, value], c[i, value + 1]
    slots = (start << 16) | (nxt - start)
    slots = torch.where(slots >= 2 ** 31, slots - 2 ** 32, slots).to(torch.int32)
    esc = torch.nonzero(neg | big).reshape(-1)
    return slots.reshape(sym.shape), esc.to(torch.int32), raw[esc]
