"""Multi-GPU plumbing: one process per GPU, the batch of images partitioned across ranks,
and ONE tiny collective per step.

The path shards by independent units (images): every element is independent given
(y, mu, sigma); the only coupling is the scalar rate/distortion sum.  So there is no
data-path collective — each rank runs the fused kernels on its images and the ranks
all-reduce a packed float64 vector {bits_y+z, squared-error sum, pixels, images}.  Compare
the reference's ``nn.DataParallel`` (src/utils/helper.py:106-113, src/train.py:168-169),
which gathers full likelihood tensors to GPU 0 every step.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _cabi


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced partition: the first n % world ranks get one extra item."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; single process if absent."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


class RateReducer:
    """Packs the per-rank scalars into one float64 vector and all-reduces it (SUM).

    ``reduce(bits, sq_err, pixels)`` returns the global {bpp, mse, bits, pixels, images};
    with world_size 1 it is a no-op wrapper, so single- and multi-GPU code paths are the same.
    """

    FIELDS = ("bits", "sq_err", "pixels", "images")

    def __init__(self, device: torch.device, slots: int = 1):
        """``slots`` > 1 batches that many consecutive steps into ONE all-reduce (a [slots, 4] matrix):
        the exchange is latency-bound, and a collective kernel that lands in the middle of a
        persistent-CTA launch costs that launch a second wave, so steps that are enqueued together
        share one collective."""
        self.device = device
        self.slots = int(slots)
        self.packed = torch.zeros(self.slots, len(self.FIELDS), dtype=torch.float64, device=device)
        self.reduced: Optional[torch.Tensor] = None

    def pack(self, bits_per_image: torch.Tensor, sq_err: Optional[torch.Tensor], pixels: float, slot: int = 0
             ) -> torch.Tensor:
        row = self.packed[slot]
        row[0] = bits_per_image.sum()
        if sq_err is not None:
            row[1] = sq_err
        else:
            row[1] = 0.0
        row[2] = float(pixels)
        row[3] = float(bits_per_image.numel())
        return self.packed

    def set_static(self, sq_err: float, pixels: float, images: float) -> None:
        """Fill the fields that do not change from step to step (once, outside the hot loop)."""
        self.packed[:, 1], self.packed[:, 2], self.packed[:, 3] = float(sq_err), float(pixels), float(images)

    def pack_bits(self, bits_per_image: torch.Tensor, slot: int = 0) -> torch.Tensor:
        """Hot-loop form: ONE tiny kernel writes sum(bits) into the packed matrix (graph-capturable)."""
        torch.sum(bits_per_image, dim=0, keepdim=True, out=self.packed[slot, 0:1])
        return self.packed

    def all_reduce(self, async_op: bool = False):
        """SUM over ranks of the packed matrix into ``self.reduced`` (the local matrix stays untouched: its static
        fields — pixels, images — are filled once and must not be summed again by the next step's collective)."""
        if self.reduced is None:
            self.reduced = torch.empty_like(self.packed)
        self.reduced.copy_(self.packed)
        if dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.reduced, op=dist.ReduceOp.SUM, async_op=async_op)
        return None

    def result(self, slot: int = 0) -> dict:
        v = (self.packed if self.reduced is None else self.reduced)[slot].tolist()
        pixels = max(v[2], 1.0)
        return {"bits": v[0], "sq_err": v[1], "pixels": v[2], "images": v[3],
                "bpp": v[0] / pixels, "mse": v[1] / pixels}


class PeerRateExchange:
    """The rate exchange WITHOUT a collective kernel (include/reslic_b200.h: reslic_rate_exchange).

    Every rank owns one small exchange buffer in peer-accessible device memory; the buffers are mapped into all
    ranks' address spaces once (CUDA IPC handles moved by ``torch.distributed``).  The Gaussian-conditional launch
    that COLLECTS a batch's rate (the last slice launch of ``TcmEntropyPath.forward(..., exchange=self)``) publishes
    the packed row {sum of bits, extra, pixels, images} into slot ``step % ring`` of EVERY rank's buffer as four
    self-validating 16-byte cells {value, step + 1} with plain stores over NVLink (no fence, no flag word); nothing else
    runs on the step — compare ``RateReducer`` (one
    NCCL all-reduce kernel per step or per graph, which lands in the middle of persistent-CTA launches) and the
    reference's ``nn.DataParallel`` gather of whole likelihood tensors (src/utils/helper.py:106-113,
    src/train.py:168-169).  ``read(n)`` adds the ``world`` rows of the next n steps in rank order on the device (a
    tiny kernel that waits on LOCAL memory only) — identical bits on every rank.

    Step numbers: a batch is published as step ``cursor + step`` — ``cursor`` a device word this object owns,
    ``step`` the batch's number relative to it, given per call (``TcmEntropyPath.forward(exchange=..., exchange_step=j)``).
    In eager use ``forward`` publishes step 0 and then advances the cursor by one (a tiny stream-ordered add).  A CUDA
    graph that holds n batches captures them with ``exchange_step=0..n-1, exchange_advance=False`` and ONE
    ``advance(n)`` behind them, so every replay publishes the next n steps; batches that run concurrently (graph
    branches, several streams) therefore land in the slot that belongs to WHICH batch they are, not to whichever
    finished first, and all ranks file the same batch under the same step.

    Discipline: every rank publishes the same sequence of steps; a rank may run at most ``ring`` steps ahead of the
    slowest reader (reads act as the back-pressure, so read at least every ``ring // 4`` steps).

    ``buffers``: single-process use (tests, one GPU standing in for several ranks) — the already-mapped bases of all
    ranks' buffers, own entry included; otherwise buffers are created and exchanged through ``group``."""

    FIELDS = ("bits", "extra", "pixels", "images")

    def __init__(self, device: torch.device, world: Optional[int] = None, rank: Optional[int] = None, ring: int = 256,
                 group=None, buffers: Optional[List[torch.Tensor]] = None):
        lib = _cabi.load()
        self.device = torch.device(device)
        self.group = group
        if buffers is not None:
            self.world, self.rank = len(buffers), int(rank or 0)
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
            if world is not None and int(world) != self.world:
                raise ValueError(f"world {world} does not match the process group ({self.world})")
        self.ring = int(ring)
        self.nbytes = int(lib.reslic_rate_exchange_bytes(self.world, self.ring))
        if self.nbytes <= 0:
            raise ValueError("bad world / ring")
        self._own_raw = None          # peer-buffer allocation of the library (multi-process form)
        self._opened: List[int] = []
        self._keep = []
        with torch.cuda.device(self.device):
            if buffers is not None:
                for b in buffers:
                    if b.device != self.device or b.numel() * b.element_size() < self.nbytes:
                        raise ValueError("exchange buffers must live on this device and hold reslic_rate_exchange_bytes()")
                bases = [int(b.data_ptr()) for b in buffers]
                self._keep = list(buffers)
            elif self.world == 1:
                own = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
                self._keep = [own]
                bases = [int(own.data_ptr())]
            else:
                bases = self._create_and_exchange(lib)
        self.bases = bases
        self.peer_table = torch.tensor(bases, dtype=torch.int64, device=self.device)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.extra = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.desc = _cabi.new(_cabi.RateExchangeDesc)
        self.desc.world, self.desc.rank, self.desc.ring = self.world, self.rank, self.ring
        self.desc.peer_base = self.peer_table.data_ptr()
        self.desc.cursor = self.cursor.data_ptr()
        self.desc.extra = self.extra.data_ptr()
        self.read_step = 0            # host-side: next step read() returns
        torch.cuda.synchronize(self.device)

    def _create_and_exchange(self, lib) -> List[int]:
        ptr = C.c_void_p()
        handle = (C.c_uint8 * _cabi.PEER_HANDLE_BYTES)()
        _cabi.check(lib.reslic_peer_buffer_create(self.nbytes, C.byref(ptr), handle), "reslic_peer_buffer_create")
        self._own_raw = int(ptr.value)
        mine = bytes(handle)
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, mine, group=self.group)
        bases = []
        for r, h in enumerate(handles):
            if r == self.rank:
                bases.append(self._own_raw)
                continue
            buf = (C.c_uint8 * _cabi.PEER_HANDLE_BYTES).from_buffer_copy(h)
            p = C.c_void_p()
            _cabi.check(lib.reslic_peer_buffer_open(buf, C.byref(p)), "reslic_peer_buffer_open")
            self._opened.append(int(p.value))
            bases.append(int(p.value))
        dist.barrier(group=self.group)        # every rank has mapped every buffer before anyone publishes
        return bases

    # ------------------------------------------------------------------ per-step static fields
    def set_static(self, pixels: float, images: float, extra: float = 0.0) -> None:
        """The row's fields that do not come from the kernels: pixels and images of this rank's share of a step, and
        ``extra`` (e.g. the squared-error sum; refill ``self.extra`` on the device before the collecting launch)."""
        self.desc.pixels, self.desc.images = float(pixels), float(images)
        self.extra.fill_(float(extra))

    def slot_of(self, step: int) -> int:
        return int(step) % self.ring

    def publish(self, bits: torch.Tensor, step: int = 0) -> None:
        """Stand-alone publisher (reslic_rate_exchange_publish_f64): the row of a batch whose per-image ``bits`` already
        exist, as step ``cursor + step``, from a one-CTA kernel on the current stream — the same row, bit for bit, that
        ``forward(..., exchange=self)`` publishes from inside the collecting launch.  In a CUDA graph put it on a branch
        of its own behind the batch's last launch: the next batch then does not wait for it."""
        if bits.dtype != torch.float64 or not bits.is_contiguous() or bits.device != self.device:
            raise ValueError("publish(): bits must be a contiguous float64 tensor on the exchange's device")
        lib = _cabi.load()
        self.desc.step = int(step)
        with torch.cuda.device(self.device):
            code = lib.reslic_rate_exchange_publish_f64(C.byref(self.desc), bits.data_ptr(), bits.numel(),
                                                        _cabi.current_stream_ptr(self.device))
        _cabi.check(code, "reslic_rate_exchange_publish_f64")

    def advance(self, n: int = 1) -> None:
        """cursor += n on the current stream (capturable): call once behind the launches that published steps
        cursor .. cursor + n - 1."""
        self.cursor.add_(int(n))

    # ------------------------------------------------------------------ reading
    def _read(self, cursor_ptr, first: int, n_steps: int, out: Optional[torch.Tensor]) -> torch.Tensor:
        n_steps = int(n_steps)
        if not 0 <= n_steps <= self.ring:
            raise ValueError("read(): between 0 and `ring` steps per call")
        if out is None:
            out = torch.empty(n_steps, 4, dtype=torch.float64, device=self.device)
        elif out.shape != (n_steps, 4) or out.dtype != torch.float64 or not out.is_contiguous() or out.device != self.device:
            raise ValueError("read(): out must be a contiguous float64 [n_steps, 4] tensor on the exchange's device")
        lib = _cabi.load()
        with torch.cuda.device(self.device):
            code = lib.reslic_rate_exchange_read_f64(self.bases[self.rank], self.world, self.ring, cursor_ptr, int(first), n_steps,
                                                     out.data_ptr(), self.status.data_ptr(),
                                                     _cabi.current_stream_ptr(self.device))
        _cabi.check(code, "reslic_rate_exchange_read_f64")
        return out

    def read(self, n_steps: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[n_steps, 4] float64 on the device: the sums over ranks of the next ``n_steps`` published steps (enqueued on
        the current stream; waits — bounded — for rows that have not arrived yet).  Step numbers are absolute and kept
        on the host, so this is the eager form; inside a CUDA graph use :meth:`read_behind`."""
        out = self._read(None, self.read_step, n_steps, out)
        self.read_step += int(n_steps)
        return out

    def read_behind(self, n_steps: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The ``n_steps`` steps published LAST (cursor - n .. cursor - 1, read from the device cursor at execution
        time): capturable, so a graph can open with a read of what its predecessor published — on a side branch it
        costs the step nothing and still keeps the ranks within a ring of each other.  Steps before 0 give zero rows."""
        return self._read(self.cursor.data_ptr(), -int(n_steps), n_steps, out)

    def skip(self, n_steps: int) -> None:
        """Advance the read position without reading (steps whose rows nobody needs)."""
        self.read_step += int(n_steps)

    def check(self) -> None:
        """Raise if a read timed out (a rank never published) or found an overwritten row (a rank ran ahead)."""
        st = int(self.status.item())
        if st:
            raise _cabi.ReslicError(f"rate exchange read failed: status {st} (1 = timeout, 2 = row overwritten)")

    @staticmethod
    def result(row) -> dict:
        v = [float(x) for x in row]
        pixels = max(v[2], 1.0)
        return {"bits": v[0], "sq_err": v[1], "pixels": v[2], "images": v[3], "bpp": v[0] / pixels, "mse": v[1] / pixels}

    def close(self) -> None:
        """Unmap the peers' buffers and free the own one (call on every rank, after a barrier)."""
        lib = _cabi.load()
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib.reslic_peer_buffer_close(p)
            self._opened = []
            if self._own_raw is not None:
                if dist.is_initialized() and self.world > 1:
                    dist.barrier(group=self.group)      # nobody still writes into the buffer being freed
                lib.reslic_peer_buffer_destroy(self._own_raw)
                self._own_raw = None
