"""The multi-GPU rate exchange without a collective kernel (include/reslic_b200.h: reslic_rate_exchange;
reslic_tcm_b200.dist.PeerRateExchange) — SURVEY.md §8e: the reference gathers whole likelihood tensors to GPU 0
(src/utils/helper.py:106-113, src/train.py:168-169) before src/training/loss.py:24-27 reduces them; here the launch
that collects a batch's rate stores one packed row into every rank's buffer.

CPU: the handle plumbing over gloo (world 2).  GPU (one device): the publish / read protocol with world 1, with two
ranks standing on one GPU (two buffers, two descriptors), ring wrap-around, CUDA-graph replays, the bounded wait.
GPU (two devices, skipped on a one-GPU box): two processes, CUDA IPC buffers, checked against dist.all_reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reslic_tcm_b200 import _cabi, synthetic
from reslic_tcm_b200 import dist as rdist

CUDA = torch.cuda.is_available()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


# ----------------------------------------------------------------------------- CPU: handle exchange over gloo
def _handle_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    rdist.init_from_env(backend="gloo")
    mine = bytes([rank + 1]) * _cabi.PEER_HANDLE_BYTES
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    torch.save(handles, os.path.join(out_dir, f"h{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_handles_are_gathered_in_rank_order(tmp_path):
    world = 2
    mp.spawn(_handle_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"h{r}.pt"))
        assert got == [bytes([q + 1]) * _cabi.PEER_HANDLE_BYTES for q in range(world)]


def test_exchange_descriptor_is_validated_without_a_gpu():
    """Argument errors are caught before any launch: wrong struct size, wrong rate mode, bad world / ring."""
    import ctypes as C

    lib = _cabi.load()
    d = _cabi.new(_cabi.GcDesc)
    d.B, d.n = 2, 8
    d.y = d.sigma = d.lik = 64            # never dereferenced: every case below fails validation first
    d.y_bs = d.sigma_bs = d.lik_bs = 8
    d.scale_bound = 0.11
    d.bits, d.workspace, d.workspace_bytes = 64, 64, 1 << 20
    x = _cabi.new(_cabi.RateExchangeDesc)
    x.world, x.rank, x.ring, x.peer_base, x.cursor, x.step = 2, 0, 8, 64, 64, 0
    d.exchange = C.pointer(x)
    d.bits_accumulate = 1                 # "+=" mode: this launch does not complete bits[]
    assert lib.reslic_gc_fwd_f32(C.byref(d), None) == -1 and b"exchange needs a rate output" in lib.reslic_last_error()
    d.bits_accumulate = _cabi.RATE_DEFERRED
    assert lib.reslic_gc_fwd_f32(C.byref(d), None) == -1
    d.bits_accumulate = _cabi.RATE_COLLECT
    x.struct_size -= 8
    assert lib.reslic_gc_fwd_f32(C.byref(d), None) == -1 and b"reslic_rate_exchange" in lib.reslic_last_error()
    x.struct_size += 8
    for field, bad in (("world", 0), ("world", 65), ("rank", 2), ("ring", 0), ("peer_base", None), ("step", -1)):
        keep = getattr(x, field)
        setattr(x, field, bad)
        assert lib.reslic_gc_fwd_f32(C.byref(d), None) == -1, field
        setattr(x, field, keep)
    assert lib.reslic_rate_exchange_read_f64(64, 2, 8, None, 0, 9, 64, 64, None) == -1      # more steps than ring slots
    assert lib.reslic_rate_exchange_read_f64(64, 2, 8, None, -1, 1, 64, 64, None) == -1     # negative step without a cursor


# ----------------------------------------------------------------------------- GPU, one device
def _path(dev):
    from reslic_tcm_b200.pipeline import TcmEntropyPath

    p = TcmEntropyPath().to(dev).eval()
    synthetic.load_eb_parameters(p.entropy_bottleneck, synthetic.eb_parameters())
    p.gaussian_conditional.scale_table = synthetic.scale_table(dev)
    return p


def _inputs(dev, images, hw=(8, 8)):
    b = synthetic.make_batch(1, images, y_hw=hw, z_hw=(hw[0] // 4, hw[1] // 4))
    return {k: b[k].to(dev) for k in ("y", "mu", "sigma", "z")}


@pytest.mark.gpu
def test_world1_publish_and_read_over_ring_wraparound_and_graph_replays():
    dev = torch.device("cuda:0")
    path, inp = _path(dev), _inputs(dev, range(5))
    ex = rdist.PeerRateExchange(dev, ring=4)
    ex.set_static(pixels=5 * 128 * 128, images=5, extra=2.5)
    want = None
    for step in range(10):                       # 2.5 times round the ring, read every step
        res = path.forward(inp["y"], inp["mu"], inp["sigma"], inp["z"], exchange=ex)
        row = ex.read(1)[0].tolist()
        bits = float(res["bits"].double().sum())
        want = want if want is not None else bits
        assert row == [bits, 2.5, 5 * 128 * 128, 5.0] and bits == want        # integer fixed-point sums: exact
    ex.check()
    assert int(ex.cursor.item()) == 10
    # graph replays publish consecutive steps (the cursor lives on the device)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        path.forward(inp["y"], inp["mu"], inp["sigma"], inp["z"], exchange=ex)
    for _ in range(3):
        g.replay()
    rows = ex.read(3)
    torch.cuda.synchronize()
    ex.check()
    assert rows[:, 0].tolist() == [want] * 3 and int(ex.cursor.item()) == 13
    # whole-y (one launch) collects and publishes as well
    path.forward(inp["y"], inp["mu"], inp["sigma"], inp["z"], exchange=ex, fuse_slices=True)
    assert ex.read(1)[0, 0].item() == pytest.approx(want, rel=1e-6)
    ex.check()


@pytest.mark.gpu
def test_concurrent_batches_of_one_graph_keep_their_slots():
    """A graph of several batches on CONCURRENT branches (bench.py's launch pattern): explicit step numbers, one advance
    behind the join — the slot of a batch must not depend on which branch finishes first."""
    dev = torch.device("cuda:0")
    ex = rdist.PeerRateExchange(dev, ring=16)
    paths = [_path(dev) for _ in range(3)]
    inps = [_inputs(dev, range(1 + k, 4 + 2 * k)) for k in range(3)]          # 3, 4, 5 images: three different rates
    for p_, i_ in zip(paths, inps):
        p_.forward(i_["y"], i_["mu"], i_["sigma"], i_["z"])
    torch.cuda.synchronize()
    wants = [float(p_.forward(i_["y"], i_["mu"], i_["sigma"], i_["z"])["bits"].double().sum()) for p_, i_ in zip(paths, inps)]
    side = [torch.cuda.Stream(dev) for _ in range(2)]
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        cur = torch.cuda.current_stream(dev)
        lanes = [cur] + side
        for s_ in side:
            s_.wait_stream(cur)
        for j in range(6):
            with torch.cuda.stream(lanes[j % 3]):
                i_ = inps[j % 3]
                ex.set_static(pixels=1, images=i_["y"].shape[0])
                paths[j % 3].forward(i_["y"], i_["mu"], i_["sigma"], i_["z"], exchange=ex, exchange_step=j, exchange_advance=False)
        for s_ in side:
            cur.wait_stream(s_)
        ex.advance(6)
    behind = torch.full((6, 4), -1.0, dtype=torch.float64, device=dev)
    g3 = torch.cuda.CUDAGraph()          # a graph that opens with a read of what was published last (capturable form)
    with torch.cuda.graph(g3):
        ex.read_behind(6, out=behind)
    g3.replay()
    torch.cuda.synchronize()
    assert behind.abs().sum().item() == 0.0                                   # nothing published yet: zero rows, no wait
    for _ in range(2):
        g2.replay()
        rows = ex.read(6)
        g3.replay()
        torch.cuda.synchronize()
        ex.check()
        assert rows[:, 0].tolist() == [wants[j % 3] for j in range(6)]
        assert rows[:, 3].tolist() == [float(inps[j % 3]["y"].shape[0]) for j in range(6)]
        assert torch.equal(behind, rows)
    assert int(ex.cursor.item()) == 12


@pytest.mark.gpu
def test_two_ranks_on_one_gpu_agree_and_match_the_unsharded_sum():
    dev = torch.device("cuda:0")
    lib = _cabi.load()
    world, ring, B = 2, 8, 7
    nbytes = lib.reslic_rate_exchange_bytes(world, ring)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    exs = [rdist.PeerRateExchange(dev, rank=r, ring=ring, buffers=bufs) for r in range(world)]
    paths = [_path(dev) for _ in range(world)]
    shards = [rdist.shard_range(B, r, world) for r in range(world)]
    local = []
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    for step in range(3):
        for r in range(world):
            inp = _inputs(dev, shards[r])
            exs[r].set_static(pixels=len(shards[r]) * 128 * 128, images=len(shards[r]), extra=float(r + 1))
            torch.cuda.synchronize()
            with torch.cuda.stream(streams[r]):      # the two "ranks" run concurrently, as two GPUs would
                res = paths[r].forward(inp["y"], inp["mu"], inp["sigma"], inp["z"], exchange=exs[r])
            if step == 0:
                torch.cuda.synchronize()
                local.append(float(res["bits"].double().sum()))
    torch.cuda.synchronize()
    rows = [ex.read(3) for ex in exs]
    torch.cuda.synchronize()
    for ex in exs:
        ex.check()
    assert torch.equal(rows[0], rows[1])                                   # every rank holds the same global rows
    total = local[0] + local[1]                                            # rank order, as the read kernel adds
    assert rows[0][0].tolist() == [total, 3.0, B * 128 * 128, float(B)]
    # ... and it is the rate of the unsharded batch (per-image sums do not depend on the sharding)
    whole = _path(dev).forward(**_inputs(dev, range(B)))
    assert float(whole["bits"].double().sum()) == pytest.approx(total, rel=1e-12)


@pytest.mark.gpu
def test_read_of_a_step_nobody_published_times_out_instead_of_hanging():
    dev = torch.device("cuda:0")
    ex = rdist.PeerRateExchange(dev, ring=4)
    out = ex.read(1)
    torch.cuda.synchronize()
    assert torch.isnan(out).all()
    with pytest.raises(_cabi.ReslicError, match="status 1"):
        ex.check()


# ----------------------------------------------------------------------------- GPU, two devices / two processes
def _peer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    rdist.init_from_env(backend="nccl")
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    B = 6
    mine = rdist.shard_range(B, rank, world)
    path, inp = _path(dev), _inputs(dev, mine)
    ex = rdist.PeerRateExchange(dev, ring=8)
    ex.set_static(pixels=len(mine) * 128 * 128, images=len(mine))
    g = torch.cuda.CUDAGraph()
    path.forward(inp["y"], inp["mu"], inp["sigma"], inp["z"])
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        res = path.forward(inp["y"], inp["mu"], inp["sigma"], inp["z"], exchange=ex)
    rows = []
    for _ in range(5):                   # 20 steps round a ring of 8, read every 4
        for _ in range(4):
            g.replay()
        rows.append(ex.read(4).clone())
    torch.cuda.synchronize()
    ex.check()
    red = rdist.RateReducer(dev)         # the checked fallback: NCCL all-reduce of the same local row
    red.pack(res["bits"], None, len(mine) * 128 * 128)
    red.all_reduce()
    torch.save({"rows": torch.cat(rows).cpu(), "nccl": red.result()}, os.path.join(out_dir, f"x{rank}.pt"))
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(not CUDA or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_processes_exchange_over_peer_memory_and_match_nccl(tmp_path):
    world = 2
    mp.spawn(_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"x{r}.pt")) for r in range(world)]
    assert torch.equal(outs[0]["rows"], outs[1]["rows"]) and outs[0]["rows"].shape == (20, 4)
    first = outs[0]["rows"][0].tolist()
    assert (outs[0]["rows"] == outs[0]["rows"][0]).all()
    assert first[0] == pytest.approx(outs[0]["nccl"]["bits"], rel=1e-12)
    assert first[2] == outs[0]["nccl"]["pixels"] == 6 * 128 * 128 and first[3] == 6.0


@pytest.mark.gpu
def test_standalone_publisher_equals_the_fused_publish():
    """reslic_rate_exchange_publish_f64 (one CTA behind the collecting launch) publishes, bit for bit, the row the
    collecting launch publishes itself."""
    dev = torch.device("cuda:0")
    path, inp = _path(dev), _inputs(dev, range(9))
    fused, branch = rdist.PeerRateExchange(dev, ring=4), rdist.PeerRateExchange(dev, ring=4)
    for ex in (fused, branch):
        ex.set_static(pixels=9 * 128 * 128, images=9, extra=0.25)
    for step in range(6):                         # round the ring
        res = path.forward(inp["y"], inp["mu"], inp["sigma"], inp["z"], exchange=fused)
        branch.publish(res["bits"], step=0)
        branch.advance(1)
        a, b = fused.read(1), branch.read(1)
        torch.cuda.synchronize()
        assert torch.equal(a, b) and a[0, 0].item() == float(res["bits"].double().sum())
    fused.check()
    branch.check()
    with pytest.raises(ValueError):
        branch.publish(res["bits"].float())
