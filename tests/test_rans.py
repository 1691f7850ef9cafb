"""CPU: host rANS coder — exact encode/decode round trips incl. the bypass escape, the
stream-continuation (decode_stream) semantics TCM.decompress relies on (tcm.py:604-621), error
behaviour, and compression close to the model entropy."""
import math

import pytest
import torch

from oracle import compressai_ref as cr
from reslic_tcm_b200 import rans


@pytest.fixture(scope="module")
def tables():
    cdf, offset, length = cr.gc_update(cr.get_scale_table())
    return cdf, length, offset


def _draw(tables, n, seed, sigma_hi=40.0):
    g = torch.Generator().manual_seed(seed)
    sigma = torch.exp(torch.empty(n).uniform_(math.log(0.05), math.log(sigma_hi), generator=g))
    sym = torch.round(sigma * torch.randn(n, generator=g)).int()
    idx = cr.build_indexes(sigma, cr.get_scale_table())
    return sym, idx, sigma


def test_round_trip_one_shot_and_list_inputs(tables):
    cdf, length, offset = tables
    sym, idx, _ = _draw(tables, 20000, 1)
    enc = rans.RansEncoder()
    s = enc.encode_with_indexes(sym.tolist(), idx.tolist(), cdf.tolist(), length.tolist(), offset.tolist())
    s2 = enc.encode_with_indexes(sym, idx, cdf, length, offset)
    assert s == s2 and isinstance(s, bytes) and len(s) % 4 == 0
    out = rans.RansDecoder().decode_with_indexes(s, idx.tolist(), cdf.tolist(), length.tolist(), offset.tolist())
    assert out == sym.tolist()


def test_bypass_escape_for_out_of_range_symbols(tables):
    cdf, length, offset = tables
    idx = torch.zeros(64, dtype=torch.int32)                      # narrowest CDF: every large value escapes
    sym = torch.tensor([0, 1, -1, 5, -5, 100, -100, 32767, -32768, 2 ** 20, -(2 ** 20), 2 ** 30, -(2 ** 30)] * 4 +
                       [0] * 12, dtype=torch.int32)
    s = rans.RansEncoder().encode_with_indexes(sym, idx, cdf, length, offset)
    out = rans.RansDecoder().decode_with_indexes(s, idx, cdf, length, offset)
    assert out == sym.tolist()


def test_buffered_encoder_and_stream_continuation(tables):
    """TCM.compress pushes all slices into one BufferedRansEncoder (tcm.py:551-565); TCM.decompress
    reads them back slice by slice from one stream (tcm.py:604-621)."""
    cdf, length, offset = tables
    parts = [_draw(tables, n, 10 + k) for k, n in enumerate((3000, 1, 0, 777, 4096))]
    enc = rans.BufferedRansEncoder()
    enc.encode_with_indexes(torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), cdf, length, offset)
    stream = enc.flush()
    dec = rans.RansDecoder()
    dec.set_stream(stream)
    for sym, idx, _ in parts:
        assert dec.decode_stream(idx, cdf, length, offset) == sym.tolist()
    # pushing in several calls gives the same stream as one call
    enc2 = rans.BufferedRansEncoder()
    for sym, idx, _ in parts:
        enc2.encode_with_indexes(sym, idx, cdf, length, offset)
    assert enc2.flush() == stream


def test_rate_is_close_to_model_entropy(tables):
    cdf, length, offset = tables
    sym, idx, sigma = _draw(tables, 200000, 3, sigma_hi=20.0)
    s = rans.RansEncoder().encode_with_indexes(sym, idx, cdf, length, offset)
    table = cr.get_scale_table()
    lik = cr.lower_bound(cr.gc_likelihood(sym.float(), table[idx.long()], None), 1e-9)   # model at the table scale
    bits_model = float(-torch.log2(lik.double()).sum())
    assert len(s) * 8 <= 1.01 * bits_model + 64


def test_errors(tables):
    cdf, length, offset = tables
    enc = rans.RansEncoder()
    with pytest.raises(ValueError):
        enc.encode_with_indexes([1, 2], [0], cdf, length, offset)
    with pytest.raises(ValueError):
        enc.encode_with_indexes([1], [64], cdf, length, offset)          # index out of range
    with pytest.raises(ValueError):
        rans.RansDecoder().set_stream(b"abc")
    dec = rans.RansDecoder()
    with pytest.raises(ValueError):
        dec.decode_stream([0], cdf, length, offset)                       # no stream
    s = enc.encode_with_indexes([3, -2, 7], [5, 5, 5], cdf, length, offset)
    with pytest.raises(ValueError):
        rans.RansDecoder().decode_with_indexes(s, [5] * 100000, cdf, length, offset)   # reads past the end


def test_entropy_model_compress_decompress_cpu_buffers(tables):
    """EntropyModel-level batch helpers (one string per image)."""
    cdf, length, offset = tables
    sym = torch.stack([_draw(tables, 512, 20 + b)[0] for b in range(3)]).reshape(3, 8, 8, 8)
    idx = torch.stack([_draw(tables, 512, 20 + b)[1] for b in range(3)]).reshape(3, 8, 8, 8)
    strings = rans.encode_with_indexes_batch(sym, idx, cdf, length, offset)
    assert len(strings) == 3
    back = rans.decode_with_indexes_batch(strings, idx, cdf, length, offset)
    assert torch.equal(back, sym)


def test_batch_helpers_thread_over_images_and_long_runs_use_the_symbol_table():
    """encode/decode_with_indexes_batch: one stream per image, images coded on a thread pool (same bytes as
    single-threaded), and the decoder's coarse symbol table (long runs) agrees with its binary search
    (short runs) — incl. bypass-coded outliers."""
    import torch
    from oracle import compressai_ref as cr
    from reslic_tcm_b200 import rans, synthetic

    table = synthetic.scale_table()
    cdf, off, cdf_len = cr.gc_update(table)
    B, n = 5, 6000                                   # > 8 * 256 symbols per image: table path
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, 64, (B, n), generator=g, dtype=torch.int32)
    sym = (torch.randn(B, n, generator=g) * table[idx.long()]).round().int()
    sym[1, :40] = torch.tensor([70000, -70000] * 20, dtype=torch.int32)
    one = rans.encode_with_indexes_batch(sym, idx, cdf, cdf_len, off, threads=1)
    many = rans.encode_with_indexes_batch(sym, idx, cdf, cdf_len, off, threads=4)
    assert one == many and len(one) == B
    for th in (1, 4):
        assert torch.equal(rans.decode_with_indexes_batch(one, idx, cdf, cdf_len, off, threads=th), sym)
    # short prefix of image 1 through the binary-search path
    short = rans.encode_with_indexes_batch(sym[1:2, :500], idx[1:2, :500], cdf, cdf_len, off)
    assert torch.equal(rans.decode_with_indexes_batch(short, idx[1:2, :500], cdf, cdf_len, off), sym[1:2, :500])


def test_push_slots_gives_the_same_stream_as_push(tables):
    """reslic_rans_encoder_push_slots (lookups done elsewhere — on the GPU by reslic_rans_slots_u32) must produce the
    byte-identical stream, escapes included."""
    from tests.util import rans_slots_reference

    cdf, length, offset = tables
    sym, idx, _ = _draw(tables, 30000, 5)
    sym[::997] = 40000                    # far outside every table: bypass escapes, also on the narrowest CDFs
    sym[5::1499] = -70000
    want = rans.RansEncoder().encode_with_indexes(sym, idx, cdf, length, offset)
    slots, esc_pos, esc_raw = rans_slots_reference(sym, idx, cdf, length, offset)
    assert esc_pos.numel() > 40
    enc = rans.BufferedRansEncoder()
    enc.encode_slots(slots, esc_pos, esc_raw)
    assert enc.flush() == want
    # two pushes, escapes given relative to each push
    half = 15000
    cut = int(torch.searchsorted(esc_pos.long(), torch.tensor([half]))[0])
    enc.encode_slots(slots[:half], esc_pos[:cut], esc_raw[:cut])
    enc.encode_slots(slots[half:], esc_pos[cut:] - half, esc_raw[cut:])
    assert enc.flush() == want
    with pytest.raises(ValueError, match="ascending"):
        enc.encode_slots(slots[:10], torch.tensor([20], dtype=torch.int32), torch.tensor([3]))
    enc.flush()


def test_reciprocal_table_equals_the_divide_for_every_frequency():
    """The encoder replaces `state / freq` by a multiply with a tabulated 64-bit reciprocal; the library's self-check
    compares the two for every frequency 1..65535 at the edges of the admissible state range and at random states."""
    from reslic_tcm_b200 import _cabi

    assert _cabi.load().reslic_rans_check_reciprocals(200, 7) == 0


def test_probability_one_symbols_are_rejected_not_miscoded():
    """A row whose only regular symbol has range 65536 does not fit the 16-bit range of a coder word."""
    cdf = torch.tensor([[0, 65536, 65536]], dtype=torch.int32)
    enc = rans.BufferedRansEncoder()
    with pytest.raises(ValueError):
        enc.encode_with_indexes([0], [0], cdf, [3], [0])
