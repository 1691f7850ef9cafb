"""CPU: the multi-GPU host logic (batch sharding + the one packed-scalar all-reduce) with
world_size 2 over gloo.  The per-image bits each rank contributes come from the oracle here
(test infrastructure); on the GPU they come from the fused kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reslic_tcm_b200 import dist as rdist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 2, 3, 7, 8, 24, 64, 255, 256):
        for world in (1, 2, 3, 4, 8):
            parts = [rdist.shard_range(n, r, world) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rdist.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = rdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    from oracle import compressai_ref as cr
    from reslic_tcm_b200 import synthetic

    mine = rdist.shard_range(B, rank, world)
    batch = synthetic.make_batch(1, mine, y_hw=(4, 4), z_hw=(1, 1))
    _, lik = cr.gc_forward(batch["y"], batch["sigma"], batch["mu"])
    bits = cr.per_image_bits(lik)
    red = rdist.RateReducer(torch.device("cpu"))
    red.pack(bits, torch.tensor(float(len(mine)) * 0.5, dtype=torch.float64), 64 * 64 * len(mine))
    red.all_reduce()
    red.all_reduce()          # a second step: the static fields (pixels, images) must not be summed over the ranks again
    res = red.result()
    torch.save({"res": res, "bits": bits, "images": list(mine)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_packed_allreduce_matches_single_process(tmp_path):
    B, world = 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert sorted(i for o in outs for i in o["images"]) == list(range(B))
    # every rank holds the same global result
    assert outs[0]["res"] == outs[1]["res"]
    # and it equals the single-process reduction over the whole batch
    from oracle import compressai_ref as cr
    from reslic_tcm_b200 import synthetic

    batch = synthetic.make_batch(1, range(B), y_hw=(4, 4), z_hw=(1, 1))
    _, lik = cr.gc_forward(batch["y"], batch["sigma"], batch["mu"])
    total = float(cr.per_image_bits(lik).sum())
    res = outs[0]["res"]
    assert res["bits"] == pytest.approx(total, rel=1e-12)
    assert res["images"] == B and res["pixels"] == 64 * 64 * B
    assert res["bpp"] == pytest.approx(total / (64 * 64 * B), rel=1e-12)
    assert res["sq_err"] == pytest.approx(0.5 * B)
    # per-image bits are unchanged by the sharding (same seeds per image)
    cat = torch.cat([o["bits"] for o in sorted(outs, key=lambda o: o["images"][0] if o["images"] else 1 << 30)])
    assert torch.equal(cat, cr.per_image_bits(lik))


def test_single_process_reducer_is_a_noop_wrapper():
    red = rdist.RateReducer(torch.device("cpu"))
    red.pack(torch.tensor([10.0, 6.0], dtype=torch.float64), None, 32)
    assert red.all_reduce() is None
    assert red.result() == {"bits": 16.0, "sq_err": 0.0, "pixels": 32.0, "images": 2.0, "bpp": 0.5, "mse": 0.0}


def _worker_slots(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    red = rdist.RateReducer(torch.device("cpu"), slots=3)
    red.set_static(0.0, 100.0, 2.0)
    for step in range(3):                       # three steps enqueued together share ONE all-reduce
        bits = torch.tensor([1.0 + rank, 10.0 * (step + 1)], dtype=torch.float64)
        red.pack_bits(bits, slot=step)
    red.all_reduce()
    torch.save([red.result(slot=k) for k in range(3)], os.path.join(out_dir, f"s{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_batched_steps_share_one_allreduce(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker_slots, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"s{r}.pt")) for r in range(world)]
    assert outs[0] == outs[1]
    for step in range(3):
        r = outs[0][step]
        assert r["bits"] == (1.0 + 10.0 * (step + 1)) + (2.0 + 10.0 * (step + 1))
        assert r["pixels"] == 200.0 and r["images"] == 4.0
        assert r["bpp"] == pytest.approx(r["bits"] / 200.0)
