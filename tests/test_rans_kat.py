"""CPU: known-answer streams for the rANS coder, derived BY HAND from the published
algorithm — ryg_rans' 64-bit coder (state x >= L = 2^31, 32-bit renormalisation words, Rans64EncPut / EncPutBits /
EncFlush) with the CompressAI framing (16-bit probabilities, value = symbol - offset, escape slot cdf_size - 2, 4-bit
bypass groups: count first, then little-endian nibbles; symbols pushed in reverse) — so that the coder is checked
against the algorithm it claims to follow and not only against its own decoder.  compressai itself is not installable
here (no network), so byte-compatibility with ITS bitstreams stays unverified; these vectors pin the arithmetic.

The arithmetic of every vector is written out in its test.  `_model_encode` is an independent transcription of the
published formulas in plain Python integers (no shared code with csrc/rans.cpp) used for the longer streams."""
import struct

import pytest
import torch

from reslic_tcm_b200 import rans

L = 1 << 31
CDF = [0, 16384, 49152, 57344, 65536]      # frequencies 1/4, 1/2, 1/8 and the escape slot 1/8
CDFS = torch.tensor([CDF], dtype=torch.int32)
SIZES = torch.tensor([5], dtype=torch.int32)      # max_value = cdf_size - 2 = 3: symbols 0..2 regular, 3 = escape
OFFS = torch.tensor([0], dtype=torch.int32)


def _words(*w):
    return struct.pack("<%dI" % len(w), *w)


def _enc(symbols, cdfs=CDFS, sizes=SIZES, offs=OFFS, indexes=None):
    idx = [0] * len(symbols) if indexes is None else indexes
    return rans.RansEncoder().encode_with_indexes(symbols, idx, cdfs, sizes, offs)


def test_three_regular_symbols():
    """symbols 0, 1, 2 — pushed in reverse, x0 = 2^31, no renormalisation (x stays below x_max = 2^47 * freq):
         2 (start 49152, freq 8192):  x = ((2^31 // 8192) << 16) + 2^31 % 8192 + 49152           = 2^34 + 49152
         1 (start 16384, freq 32768): (2^34 + 49152) // 32768 = 2^19 + 1 rem 16384;
                                      x = ((2^19 + 1) << 16) + 16384 + 16384                      = 2^35 + 98304
         0 (start 0, freq 16384):     (2^35 + 98304) // 16384 = 2^21 + 6 rem 0;  x = (2^21 + 6) << 16 = 2^37 + 393216
       flush: low word 393216 = 0x00060000, high word 2^37 >> 32 = 0x20."""
    got = _enc([0, 1, 2])
    assert got == _words(0x00060000, 0x00000020) == bytes.fromhex("0000060020000000")
    assert rans.RansDecoder().decode_with_indexes(got, [0, 0, 0], CDFS, SIZES, OFFS) == [0, 1, 2]


def test_one_escape_with_one_bypass_nibble():
    """symbols 1, 5: value 5 >= max_value 3 -> raw = 2 * (5 - 3) = 4, coded as the escape slot (start 57344, freq 8192)
       followed by the bypass groups: count 1, then the nibble 4.  Reverse order, EncPutBits(x, v, 4): x = (x << 4) | v:
         nibble 4: x = 2^35 + 4
         count 1:  x = (2^35 + 4) << 4 | 1                                  = 2^39 + 65
         escape:   (2^39 + 65) // 8192 = 2^26 rem 65;  x = (2^26 << 16) + 65 + 57344               = 2^42 + 57409
         1:        (2^42 + 57409) // 32768 = 2^27 + 1 rem 24641; x = ((2^27 + 1) << 16) + 24641 + 16384 = 2^43 + 106561
       flush: low word 106561 = 0x0001A041, high word 2^43 >> 32 = 0x800."""
    got = _enc([1, 5])
    assert got == _words(0x0001A041, 0x00000800) == bytes.fromhex("41a0010000080000")
    assert rans.RansDecoder().decode_with_indexes(got, [0, 0], CDFS, SIZES, OFFS) == [1, 5]


def test_negative_value_escape_and_offset():
    """offset -2, symbol -4: value = -4 - (-2) = -2 < 0 -> raw = -2 * value - 1 = 3 (odd = negative side), escape slot,
       count 1, nibble 3:   x = 2^35 + 3;  x = (2^35 + 3) << 4 | 1 = 2^39 + 49;  escape: 2^39 + 49 = 8192 * 2^26 + 49,
       x = (2^26 << 16) + 49 + 57344 = 2^42 + 57393  -> words 57393 = 0x0000E031, 2^42 >> 32 = 0x400."""
    offs = torch.tensor([-2], dtype=torch.int32)
    got = _enc([-4], offs=offs)
    assert got == _words(0x0000E031, 0x00000400)
    assert rans.RansDecoder().decode_with_indexes(got, [0], CDFS, SIZES, offs) == [-4]


# ----------------------------------------------------------------------------- independent model for long streams
def _model_encode(symbols, indexes, cdfs, sizes, offsets):
    ops = []
    for s, ci in zip(symbols, indexes):
        cdf, max_value = cdfs[ci], sizes[ci] - 2
        v = s - offsets[ci]
        raw = None
        if v < 0:
            raw, v = -2 * v - 1, max_value
        elif v >= max_value:
            raw, v = 2 * (v - max_value), max_value
        ops.append(("sym", cdf[v], cdf[v + 1] - cdf[v]))
        if raw is not None:
            n = 0
            while raw >> (4 * n):
                n += 1
            val = n
            while val >= 15:
                ops.append(("bits", 15))
                val -= 15
            ops.append(("bits", val))
            ops += [("bits", (raw >> (4 * j)) & 15) for j in range(n)]
    x, out = L, []
    for op in reversed(ops):
        if op[0] == "sym":
            _, start, freq = op
            if x >= ((L >> 16) << 32) * freq:
                out.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x // freq) << 16) + (x % freq) + start
        else:
            if x >= ((L >> 16) << 32) * (1 << 12):
                out.append(x & 0xFFFFFFFF)
                x >>= 32
            x = (x << 4) | op[1]
    out += [x >> 32, x & 0xFFFFFFFF]       # emitted backwards: the decoder reads state low, state high, then the rest
    return _words(*reversed(out))


def test_model_reproduces_the_hand_derived_vectors():
    assert _model_encode([0, 1, 2], [0, 0, 0], [CDF], [5], [0]) == bytes.fromhex("0000060020000000")
    assert _model_encode([1, 5], [0, 0], [CDF], [5], [0]) == bytes.fromhex("41a0010000080000")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_long_streams_with_renormalisation_escapes_and_many_tables(seed):
    g = torch.Generator().manual_seed(seed)
    n_cdfs, width = 7, 23
    rows, sizes = [], []
    for i in range(n_cdfs):
        k = 3 + int(torch.randint(0, width - 3, (1,), generator=g))          # cdf entries of this row
        f = torch.randint(1, 2000, (k - 1,), generator=g).double()
        f = torch.floor(f / f.sum() * (65536 - (k - 1))).long() + 1          # every frequency >= 1
        f[0] += 65536 - int(f.sum())
        rows.append([0] + torch.cumsum(f, 0).tolist() + [0] * (width - k))
        sizes.append(k)
    offsets = torch.randint(-6, 3, (n_cdfs,), generator=g).tolist()
    n = 5000
    indexes = torch.randint(0, n_cdfs, (n,), generator=g).tolist()
    symbols = torch.randint(-40, 60, (n,), generator=g).tolist()
    symbols[17], symbols[18] = 10 ** 6, -10 ** 6                             # escapes of 6 nibbles
    symbols[19] = 2 ** 30                                                    # 8 nibbles
    want = _model_encode(symbols, indexes, rows, sizes, offsets)
    cd, sz, of = torch.tensor(rows, dtype=torch.int32), torch.tensor(sizes, dtype=torch.int32), torch.tensor(offsets, dtype=torch.int32)
    got = rans.RansEncoder().encode_with_indexes(symbols, indexes, cd, sz, of)
    assert len(want) > 2000 and got == want                                  # hundreds of renormalisation words
    assert rans.RansDecoder().decode_with_indexes(got, indexes, cd, sz, of) == symbols
