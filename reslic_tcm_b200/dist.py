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

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


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
        if dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.packed, op=dist.ReduceOp.SUM, async_op=async_op)
        return None

    def result(self, slot: int = 0) -> dict:
        v = self.packed[slot].tolist()
        pixels = max(v[2], 1.0)
        return {"bits": v[0], "sq_err": v[1], "pixels": v[2], "images": v[3],
                "bpp": v[0] / pixels, "mse": v[1] / pixels}
