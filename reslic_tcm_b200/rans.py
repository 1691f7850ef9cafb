"""rANS coder front-end: the API of ``compressai.ans`` (RansEncoder, BufferedRansEncoder,
RansDecoder — reference call sites src/models/reference/tcm.py:2,522,564-565,604-605,621 and
src/entropy_models/coder.py:26-29) over the host C++ coder in ``csrc/rans.cpp``.

Symbols and indexes may be passed as Python lists (the reference's ``.tolist()`` habit,
tcm.py:551-552) or — cheaper — as int32 tensors straight from the fused kernel; GPU tensors are
brought to the host with one copy instead of a per-element Python list.
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Union

import torch
from torch import Tensor

from . import _cabi

IntSeq = Union[Sequence[int], Tensor]


def _i32(x: IntSeq) -> Tensor:
    if isinstance(x, Tensor):
        return x.detach().reshape(-1).to(device="cpu", dtype=torch.int32).contiguous()
    return torch.tensor(list(x), dtype=torch.int32).reshape(-1)


class _Tables:
    """cdfs [n, stride], cdf sizes [n], offsets [n] as contiguous host int32."""

    def __init__(self, cdfs, cdfs_sizes, offsets):
        if isinstance(cdfs, Tensor):
            c = cdfs.detach().to(device="cpu", dtype=torch.int32)
        else:
            width = max(len(r) for r in cdfs)
            c = torch.zeros(len(cdfs), width, dtype=torch.int32)
            for i, r in enumerate(cdfs):
                c[i, : len(r)] = torch.tensor(list(r), dtype=torch.int32)
        if c.dim() != 2:
            raise ValueError("cdfs must be 2-D")
        self.cdfs = c.contiguous()
        self.sizes = _i32(cdfs_sizes)
        self.offsets = _i32(offsets)
        if self.sizes.numel() != c.shape[0] or self.offsets.numel() != c.shape[0]:
            raise ValueError("cdfs, cdfs_sizes and offsets disagree on the number of CDFs")

    def args(self):
        return (self.cdfs.data_ptr(), self.cdfs.shape[0], self.cdfs.shape[1], self.sizes.data_ptr(),
                self.offsets.data_ptr())


class BufferedRansEncoder:
    """compressai.ans.BufferedRansEncoder: accumulate encode_with_indexes calls, flush() once."""

    def __init__(self):
        self._lib = _cabi.load()
        self._h = self._lib.reslic_rans_encoder_create()
        if not self._h:
            raise MemoryError("rans encoder")

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.reslic_rans_encoder_destroy(self._h)
            self._h = None

    def encode_with_indexes(self, symbols: IntSeq, indexes: IntSeq, cdfs, cdfs_sizes, offsets) -> None:
        s, i = _i32(symbols), _i32(indexes)
        if s.numel() != i.numel():
            raise ValueError("symbols and indexes must have the same length")
        t = cdfs if isinstance(cdfs, _Tables) else _Tables(cdfs, cdfs_sizes, offsets)
        code = self._lib.reslic_rans_encoder_push(self._h, s.data_ptr(), i.data_ptr(), s.numel(), *t.args())
        if code != 0:
            raise ValueError(self._lib.reslic_last_error().decode())

    def encode_slots(self, slots: Tensor, esc_pos: Optional[Tensor] = None, esc_raw: Optional[Tensor] = None) -> None:
        """Symbols whose table lookups ran on the device (ops.rans_slots): host int32 ``slots`` and the escapes
        that belong to them, positions ascending and relative to ``slots[0]``."""
        s = slots.detach().reshape(-1).to(device="cpu", dtype=torch.int32).contiguous()
        n_esc = 0 if esc_pos is None else esc_pos.numel()
        ep = er = None
        if n_esc:
            ep = esc_pos.detach().reshape(-1).to(device="cpu", dtype=torch.int32).contiguous()
            er = esc_raw.detach().reshape(-1).to(device="cpu", dtype=torch.int64).contiguous()
        code = self._lib.reslic_rans_encoder_push_slots(self._h, s.data_ptr(), s.numel(), ep.data_ptr() if n_esc else None,
                                                        er.data_ptr() if n_esc else None, n_esc)
        if code != 0:
            raise ValueError(self._lib.reslic_last_error().decode())

    def flush(self) -> bytes:
        data = C.c_void_p()
        n = self._lib.reslic_rans_encoder_flush(self._h, C.byref(data))
        if n < 0:
            raise RuntimeError(self._lib.reslic_last_error().decode())
        return C.string_at(data, n)


class RansEncoder:
    """compressai.ans.RansEncoder: one-shot encode_with_indexes -> bytes."""

    def encode_with_indexes(self, symbols: IntSeq, indexes: IntSeq, cdfs, cdfs_sizes, offsets) -> bytes:
        enc = BufferedRansEncoder()
        enc.encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets)
        return enc.flush()


class RansDecoder:
    """compressai.ans.RansDecoder: decode_with_indexes (one shot) or set_stream + decode_stream."""

    def __init__(self):
        self._lib = _cabi.load()
        self._h = None

    def __del__(self):
        self._close()

    def _close(self):
        if getattr(self, "_h", None):
            self._lib.reslic_rans_decoder_destroy(self._h)
            self._h = None

    def set_stream(self, encoded: bytes) -> None:
        self._close()
        self._buf = bytes(encoded)
        self._h = self._lib.reslic_rans_decoder_create(self._buf, len(self._buf))
        if not self._h:
            raise ValueError(self._lib.reslic_last_error().decode())

    def decode_stream_tensor(self, indexes: IntSeq, cdfs, cdfs_sizes, offsets) -> Tensor:
        if not self._h:
            raise ValueError("no stream set")
        i = _i32(indexes)
        t = cdfs if isinstance(cdfs, _Tables) else _Tables(cdfs, cdfs_sizes, offsets)
        out = torch.empty(i.numel(), dtype=torch.int32)
        code = self._lib.reslic_rans_decoder_decode(self._h, i.data_ptr(), i.numel(), *t.args(), out.data_ptr())
        if code != 0:
            raise ValueError(self._lib.reslic_last_error().decode())
        return out

    def decode_stream(self, indexes: IntSeq, cdfs, cdfs_sizes, offsets) -> List[int]:
        return self.decode_stream_tensor(indexes, cdfs, cdfs_sizes, offsets).tolist()

    def decode_with_indexes(self, encoded: bytes, indexes: IntSeq, cdfs, cdfs_sizes, offsets) -> List[int]:
        self.set_stream(encoded)
        return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets)


# ---- batched helpers used by EntropyModel.compress / decompress (one string per image)
def _pool_map(fn, n: int, threads: Optional[int]):
    """fn(0..n-1) on a thread pool: the coder calls are ctypes foreign calls (the GIL is released), and
    every image has its own rANS state, so images code in parallel."""
    workers = min(n, threads if threads else (os.cpu_count() or 1))
    if workers <= 1:
        return [fn(b) for b in range(n)]
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(fn, range(n)))


def encode_with_indexes_batch(symbols: Tensor, indexes: Tensor, cdf: Tensor, cdf_length: Tensor, offset: Tensor,
                              threads: Optional[int] = None) -> List[bytes]:
    """One rANS string per image (tcm.py:551-565 builds exactly these, image by image, from Python lists)."""
    t = _Tables(cdf, cdf_length.reshape(-1), offset.reshape(-1))
    s = symbols.detach().to(device="cpu", dtype=torch.int32)
    i = indexes.detach().to(device="cpu", dtype=torch.int32)

    def one(b):
        enc = BufferedRansEncoder()
        enc.encode_with_indexes(s[b], i[b], t, None, None)
        return enc.flush()

    return _pool_map(one, s.shape[0], threads)


def encode_slots_batch(slots: Tensor, esc_pos: Tensor, esc_raw: Tensor, status: Tensor,
                       threads: Optional[int] = None) -> Optional[List[bytes]]:
    """One rANS string per image from the output of ops.rans_slots over a [B, ...] batch: one D2H copy of the packed
    slots (4 bytes per symbol instead of 8), the escape list sorted and cut per image on the host.  Returns None
    when the escape list overflowed its capacity (the caller then codes from symbols and indexes); raises on
    invalid tables or indexes."""
    st = status.detach().cpu()
    n_esc, err = int(st[0]), int(st[1])
    if err:
        raise ValueError("rans_slots: " + ("cdf index out of range" if err & 1 else "invalid cdf (zero or negative frequency)"))
    if n_esc > esc_pos.numel():
        return None
    s = slots.detach().to(device="cpu")
    B = s.shape[0]
    per = s[0].numel() if B else 0
    pos = esc_pos[:n_esc].detach().cpu()
    raw = esc_raw[:n_esc].detach().cpu()
    if n_esc:
        pos, order = torch.sort(pos)
        raw = raw[order]
    cuts = torch.searchsorted(pos.to(torch.int64), torch.arange(B + 1, dtype=torch.int64) * per).tolist()

    def one(b):
        enc = BufferedRansEncoder()
        lo, hi = cuts[b], cuts[b + 1]
        enc.encode_slots(s[b], (pos[lo:hi] - b * per) if hi > lo else None, raw[lo:hi] if hi > lo else None)
        return enc.flush()

    return _pool_map(one, B, threads)


def decode_with_indexes_batch(strings: Sequence[bytes], indexes: Tensor, cdf: Tensor, cdf_length: Tensor,
                              offset: Tensor, threads: Optional[int] = None) -> Tensor:
    t = _Tables(cdf, cdf_length.reshape(-1), offset.reshape(-1))
    i = indexes.detach().to(device="cpu", dtype=torch.int32)
    out = torch.empty(i.shape, dtype=torch.int32)

    def one(b):
        dec = RansDecoder()
        dec.set_stream(strings[b])
        out[b] = dec.decode_stream_tensor(i[b], t, None, None).reshape(i[b].shape)

    _pool_map(one, len(strings), threads)
    return out.to(indexes.device)
