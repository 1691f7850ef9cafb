#!/usr/bin/env python
"""bench.py — entropy-model hot path throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic latents of the named
config (default: BASELINE.json configs[1] = TCM N=64, Kodak-shaped 768x512, batch 24,
eval-mode round quantization + build_indexes): 1 EntropyBottleneck launch on z + 5
Gaussian-conditional launches (one per channel slice, as TCM drives them) + the per-image
rate sum.  Prints ONE JSON line (rank 0).

* value      — latent elements (y + z) of ALL ranks / second, inputs resident in HBM,
               steps replayed as CUDA graphs, CUDA-event timed, max over ranks.
* e2e        — same metric through the public API with HOST (pinned) buffers: H2D of
               y/mu/sigma/z and D2H of symbols/indexes/bits inside the timed region.
* roofline   — the Gaussian-conditional kernel alone: algorithmic bytes / its average
               launch duration (graph of the 5 slice launches replayed back to back).
* cpu_baseline / --impl reference — the oracle restatement of the reference's own CPU op
               chain (oracle/compressai_ref.py) on the host cores, bounded sample.
Multi-GPU: weak scaling — every rank processes its own config-shaped batch of distinct
images; the only collective is one packed-scalar all-reduce per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from reslic_tcm_b200 import ops, synthetic  # noqa: E402

METRIC = "entropy_model_latent_melem_per_s"
UNIT = "Melem/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", type=int, default=2, choices=sorted(synthetic.CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nbuf", type=int, default=3, help="rotating buffer sets (L2-cold inputs)")
    ap.add_argument("--cpu-images", type=int, default=0, help="images in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-whole-y", action="store_true")
    ap.add_argument("--no-training-kernels", action="store_true")
    ap.add_argument("--rotations-per-graph", type=int, default=2,
                    help="consecutive steps chained in one CUDA graph = this many rotations of the buffer sets (0: one graph per step)")
    ap.add_argument("--e2e-chunks", type=int, default=1,
                    help="image chunks per batch in the host pipeline (batches are double-buffered, so 1 streams best; more chunks cut per-batch latency)")
    return ap.parse_args()


def bytes_per_y_elem(c: synthetic.Config) -> int:
    """Algorithmic HBM bytes per y element of one GC launch: read y, mu,
    sigma (12) + write y_hat, L (8) [+ symbols, indexes (8)] [+ noisy y (4) in training]."""
    return 12 + 8 + (8 if c.with_indexes else 0) + (4 if c.training else 0)


def ref_eb(params):
    from oracle import compressai_ref as cr

    eb = cr.EntropyBottleneckRef(synthetic.Z_CHANNELS)
    eb.matrices = [params[f"_matrix{i}"] for i in range(5)]
    eb.biases = [params[f"_bias{i}"] for i in range(5)]
    eb.factors = [params[f"_factor{i}"] for i in range(4)]
    eb.quantiles = params["quantiles"]
    return eb


# --------------------------------------------------------------------------- CPU arm
def cpu_reference_throughput(c: synthetic.Config, n_images: int, repeats: int, budget_s: float = 25.0):
    """Time the oracle (reference op order, unfused torch CPU ops, all host threads) on
    `n_images` images of the config.  Returns (Melem/s best, seconds per pass list, cores)."""
    from oracle import compressai_ref as cr

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synthetic.make_batch(c.cfg, range(n_images), with_noise=c.training)
    eb = ref_eb(synthetic.eb_parameters())
    table = synthetic.scale_table()
    elems = n_images * (c.y_elems_per_image + c.z_elems_per_image)
    times = []
    t_all = time.perf_counter()
    for i in range(repeats + 1):
        t0 = time.perf_counter()
        with torch.no_grad():
            cr.tcm_entropy_step(batch["y"], batch["mu"], batch["sigma"], batch["z"], eb, table,
                                training=c.training, with_indexes=c.with_indexes,
                                num_pixels=n_images * c.num_pixels_per_image,
                                noise_y=batch.get("noise_y"), noise_z=batch.get("noise_z"))
        dt = time.perf_counter() - t0
        if i > 0:  # first pass is the warm-up
            times.append(dt)
        if time.perf_counter() - t_all > budget_s and times:
            break
    best = min(times)
    return elems / best / 1e6, times, cores, elems


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port;
    compressai itself is not installable here), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = synthetic.CONFIGS[args.config]
    n_img = args.cpu_images or min(c.batch, 8)
    from oracle import compressai_ref as cr

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synthetic.make_batch(c.cfg, range(n_img), with_noise=c.training)
    eb = ref_eb(synthetic.eb_parameters())
    table = synthetic.scale_table()
    elems = n_img * (c.y_elems_per_image + c.z_elems_per_image)

    def step():
        with torch.no_grad():
            cr.tcm_entropy_step(batch["y"], batch["mu"], batch["sigma"], batch["z"], eb, table,
                                training=c.training, with_indexes=c.with_indexes,
                                num_pixels=n_img * c.num_pixels_per_image,
                                noise_y=batch.get("noise_y"), noise_z=batch.get("noise_z"))

    # K steps as asked, W warm-up steps, but bounded in time (a slow host must not turn the arm into a long job):
    # at most ~20 s of warm-up and ~120 s of timed steps; `steps` / `warmup` in the line are what actually ran
    want_steps, want_warm = max(1, args.steps), max(1, args.warmup)
    warm, t0 = 0, time.perf_counter()
    while warm < want_warm and (warm == 0 or time.perf_counter() - t0 < 20.0):
        step()
        warm += 1
    steps, t0 = 0, time.perf_counter()
    while steps < want_steps and (steps == 0 or time.perf_counter() - t0 < 120.0):
        step()
        steps += 1
    dt = time.perf_counter() - t0
    val = elems * steps / dt / 1e6
    sample = f"{n_img} of {c.batch} images of config {c.cfg} per step ({elems} latent elements), {steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": c.name, "cfg": c.cfg, "images_per_step": n_img, "cpu": cpu_model()},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed loops run."""

    def __init__(self, index: int, uuid: str = None, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.uuid, self.period = index, uuid, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.active = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(self.uuid)
                except Exception:
                    try:
                        h = nv.nvmlDeviceGetHandleByUUID(self.uuid.encode())
                    except Exception:
                        h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop.is_set():
                if self.active.is_set():
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    mask = get_reasons(h)
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                time.sleep(self.period)
        except Exception as exc:  # NVML missing: report, do not fake
            self.error = repr(exc)

    def stop(self):
        self._stop.set()

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        out = {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}
        if self.error:
            out["error"] = self.error
        return out


# --------------------------------------------------------------------------- GPU arm
def bind_to_gpu_numa(dev) -> None:
    """Multi-rank runs: pin this process to the CPUs local to its GPU (sysfs local_cpulist) so that its
    pinned host buffers are allocated on that NUMA node and the e2e copies do not cross sockets."""
    try:
        p = torch.cuda.get_device_properties(dev)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            print(f"[bench] rank on {bdf}: bound to {len(cpus)} local CPUs ({txt})", file=sys.stderr)
    except Exception as e:      # topology not exposed in this container: run unbound
        print(f"[bench] NUMA binding skipped: {e}", file=sys.stderr)


def run_ours(args):
    import torch.distributed as dist

    from reslic_tcm_b200 import _cabi, dist as rdist
    from reslic_tcm_b200.pipeline import TcmEntropyPath

    _cabi.load()  # no extension -> loud failure, never a fallback
    # stdout carries exactly ONE JSON line: libraries that write banners to fd 1 (NCCL prints its version
    # there at the first collective) are sent to stderr until that line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, local, world = rdist.init_from_env()
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    assert torch.cuda.is_available(), "bench.py (ours) needs a CUDA device"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        bind_to_gpu_numa(dev)      # before any pinned allocation: first touch puts the pages on the GPU's node
    c = synthetic.CONFIGS[args.config]
    B = c.batch                                   # weak scaling: every rank gets a full batch
    images = range(rank * B, (rank + 1) * B)      # of distinct images
    host = synthetic.make_batch(c.cfg, images, with_noise=False, pin=True)
    params = synthetic.eb_parameters()
    kw = dict(training=c.training, with_indexes=c.with_indexes, num_pixels=c.num_pixels_per_image, seed=1234)

    sets = []
    for i in range(max(1, args.nbuf)):
        path = TcmEntropyPath().to(dev).eval()
        synthetic.load_eb_parameters(path.entropy_bottleneck, params)
        path.gaussian_conditional.scale_table = synthetic.scale_table(dev)
        inp = {k: host[k].to(dev, non_blocking=True) for k in ("y", "mu", "sigma", "z")}
        torch.cuda.synchronize()
        graph, res = path.capture(inp["y"], inp["mu"], inp["sigma"], inp["z"], **kw)
        sets.append({"path": path, "inp": inp, "graph": graph, "res": res})
    # rate exchange: steps that are enqueued together (one graph of `group` steps) share ONE packed
    # all-reduce of a [group, 4] float64 matrix; two reducers alternate so that group g+1 never waits
    # on group g's collective
    group = len(sets) * max(1, args.rotations_per_graph)
    reducers = [rdist.RateReducer(dev, slots=group) for _ in range(2)]
    single = rdist.RateReducer(dev)
    for r_ in reducers + [single]:
        r_.set_static(0.0, B * c.num_pixels_per_image, B)
    y_elems, z_elems = B * c.y_elems_per_image, B * c.z_elems_per_image
    elems_rank = y_elems + z_elems
    launches_per_step = 1 + synthetic.NUM_SLICES

    # One graph per step (one batch), plus graphs of `group` consecutive steps — one per buffer set — so
    # that back-to-back batches are chained by programmatic (PDL) edges instead of graph-replay boundaries;
    # the group graphs (one per reducer) end with the `group` tiny kernels that pack sum(bits) of each step.
    super_graphs = []
    if group > 1 and args.rotations_per_graph > 0:
        for red in reducers:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for j in range(group):
                    s = sets[j % len(sets)]
                    s["path"].forward(s["inp"]["y"], s["inp"]["mu"], s["inp"]["sigma"], s["inp"]["z"], **kw)
                    if world > 1:
                        red.pack_bits(s["res"]["bits"], slot=j)
            super_graphs.append(g)

    works = [None, None, None]

    def step(i, collective=True):
        sets[i % len(sets)]["graph"].replay()
        if world > 1 and collective:  # the path's only exchange: one packed-scalar all-reduce
            if works[2] is not None:
                works[2].wait()
            single.pack_bits(sets[i % len(sets)]["res"]["bits"])
            works[2] = single.all_reduce(async_op=True)

    def run_steps(i0, n, collective=True):
        """Steps i0 .. i0+n-1, in groups of `group` where they line up with the buffer rotation."""
        i, end = i0, i0 + n
        while i < end:
            if super_graphs and i % group == 0 and i + group <= end:
                k = (i // group) % 2
                if works[k] is not None:       # stream-level wait: this reducer's previous collective has read its matrix
                    works[k].wait()
                    works[k] = None
                super_graphs[k].replay()
                if world > 1 and collective:
                    works[k] = reducers[k].all_reduce(async_op=True)
                i += group
            else:
                step(i, collective)
                i += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    sampler.start()
    warm = max(args.warmup, 3)
    run_steps(0, warm + (-warm) % group)        # warm-up ends on a rotation boundary
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    e0.record()
    run_steps(0, args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # the timed region may be shorter than an NVML sample: keep the identical loop running
    # (untimed) until the sampler has seen >= 1 s of it
    t_end = time.perf_counter() + max(0.0, 1.0 - ms / 1e3)
    i = args.steps
    i += (-i) % group
    while time.perf_counter() < t_end:
        run_steps(i, group, collective=False)   # time-bounded loop: ranks run different counts, so no collectives here
        i += group
        if i % (64 * group) < group:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    sampler.active.clear()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = elems_rank * world / (ms_per_step * 1e-3) / 1e6

    # ---- roofline leg: the GC kernel alone, 5 slice launches per graph, replayed back to back
    bpe = bytes_per_y_elem(c)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def gc_only_leg(fuse):
        """Only the GC launches of a step (5 per-slice, or 1 over the whole y) over the rotating buffer
        sets, captured as ONE graph of >= 60 launches so that the kernel's launch duration is not diluted
        by graph-replay boundaries (inside a graph consecutive launches are PDL edges); the rate stays deferred
        in the workspace and is drained after the timed region.  Returns us per launch."""
        n_launch = 1 if fuse else synthetic.NUM_SLICES
        kk = dict(kw, skip_z=True, fuse_slices=fuse, defer_rate=True)   # the kernel alone: rate finalised once per graph
        passes = max(1, -(-60 // (n_launch * len(sets))))

        def run_all():
            for _ in range(passes):
                for s in sets:
                    s["path"].forward(s["inp"]["y"], s["inp"]["mu"], s["inp"]["sigma"], s["inp"]["z"], **kk)

        def drain():   # the deferred sums (48.16 fixed point in 64 bits: no overflow over any run length) -> bits
            for s in sets:
                b = s["path"].buffers(s["inp"]["y"], s["inp"]["z"], kw.get("with_indexes", False), kw.get("training", False))
                ops.rate_finalize(b["workspace"], s["inp"]["y"].shape[0], bits=b["bits"])

        run_all()
        drain()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run_all()
        per_graph = passes * len(sets) * n_launch
        reps = max(3, -(-max(args.steps, 50) * n_launch // per_graph))
        for i in range(2):
            g.replay()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for i in range(reps):
            g.replay()
        r1.record()
        drain()                # outside the timed region: this leg times the kernel alone
        torch.cuda.synchronize()
        return r0.elapsed_time(r1) * 1e3 / (reps * per_graph), n_launch

    traffic_db = {}
    try:
        traffic_db = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get("gc_fwd_kernel", {})
    except Exception:
        pass

    def roof_obj(us, n_launch):
        elems = y_elems // n_launch
        achieved = bpe * elems / (us * 1e-6) / 1e9
        traffic = None      # dram read+write bytes per launch from the committed ncu --set full captures
        for t in (traffic_db if isinstance(traffic_db, list) else [traffic_db]):
            if t.get("config") == c.cfg and t.get("elems_per_launch") == elems:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "gc_fwd_kernel", "bytes_per_elem": bpe, "elems_per_launch": elems,
                "us_per_launch": us, "peak_source": peak_src, "gc_melem_per_s": elems / us}

    us_slice, n5 = gc_only_leg(False)
    roof = roof_obj(us_slice, n5)
    roof["launch"] = "one 64-channel slice of y per launch (TCM's call pattern, tcm.py:443-457)"

    # ---- whole-y mode: all 320 channels in ONE GC launch (models whose mu/sigma exist for all
    # channels at once, e.g. the reference's ScaleHyperprior); reported beside the per-slice mode
    whole = None
    if not args.no_whole_y:
        wgraphs = []
        for s in sets:
            kk = dict(kw, fuse_slices=True)
            s["path"].forward(s["inp"]["y"], s["inp"]["mu"], s["inp"]["sigma"], s["inp"]["z"], **kk)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                s["path"].forward(s["inp"]["y"], s["inp"]["mu"], s["inp"]["sigma"], s["inp"]["z"], **kk)
            wgraphs.append(g)
        for i in range(5):
            wgraphs[i % len(wgraphs)].replay()
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for i in range(args.steps):
            wgraphs[i % len(wgraphs)].replay()
        w1.record()
        barrier()
        wms = w0.elapsed_time(w1)
        if world > 1:
            tt = torch.tensor([wms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            wms = float(tt.item())
        us_whole, n1 = gc_only_leg(True)
        whole = {"value": elems_rank * world / (wms / args.steps * 1e-3) / 1e6, "unit": UNIT,
                 "ms_per_step": wms / args.steps, "launches_per_step": 2, "roofline": roof_obj(us_whole, n1)}

    # ---- e2e leg: public API with host buffers (H2D inputs, D2H symbols/indexes/bits)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, c, sets[0], host, dev, world, B, elems_rank, kw)

    # ---- training-path kernels on a config-5 slice (single-GPU runs): backward, noise-mode bottleneck, STanH
    training_kernels = None
    if world == 1 and not args.no_training_kernels:
        training_kernels = training_kernels_leg(dev, peak)

    # ---- e2e with the rANS table lookup on the device: one packed (start, range) slot per symbol comes back instead
    # of int32 symbols + int32 indexes (what the host coder consumes either way) — reported beside `e2e`
    e2e_slots = None
    if not args.no_e2e and c.with_indexes:
        gcm = sets[0]["path"].gaussian_conditional
        gcm.update()                                   # CDF tables of the scale table in place (setup, untimed)
        e2e_slots = run_e2e(args, c, sets[0], host, dev, world, B, elems_rank, kw, packed_slots=True)

    sampler.stop()
    clocks = sampler.summary()

    # ---- CPU baseline (rank 0, single-GPU runs only): oracle on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_img = args.cpu_images or min(c.batch, 8)
        v, times, cores, el = cpu_reference_throughput(c, n_img, repeats=8, budget_s=20.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_img} of {c.batch} images of config {c.cfg} ({el} latent elements), best of {len(times)} passes "
                         f"after 1 warm-up, oracle/compressai_ref.py tcm_entropy_step, torch {torch.__version__} CPU",
               "cpu": cpu_model()}

    if rank == 0:
        bits = sets[0]["res"]["bits"].double().cpu()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c.name, "cfg": c.cfg, "images_per_gpu": B, "y_shape": [B, 320, *c.y_hw],
                       "z_shape": [B, 192, *c.z_hw], "launches_per_step": launches_per_step,
                       "mode": f"per-slice launches (1 EB + 5 GC per step) replayed as CUDA graphs, steps chained in graphs of {group} (one packed rate all-reduce per graph when N > 1)",
                       "l2": f"{len(sets)} rotating buffer sets of {(bpe * y_elems + 12 * z_elems) / 1e6:.0f} MB each (> 126 MB L2)",
                       "bpp_mean_image0_set": float(bits.mean()) / c.num_pixels_per_image},
            "roofline": roof, "whole_y": whole, "training_kernels": training_kernels, "cpu_baseline": cpu, "e2e": e2e, "e2e_slots": e2e_slots,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def training_kernels_leg(dev, peak):
    """Per-launch times of the training-path kernels on one config-5 slice (256 patches of 256x256: y slice
    [256, 64, 16, 16], z [256, 192, 4, 4]) — the Gaussian-conditional backward, the noise-mode bottleneck
    forward / backward and the STanH family (soft beta = 10, hard, backward, compute_gap).  Each op is
    captured as a graph of 12 launches over 3 rotating input sets (403 MB > L2 for the y-shaped ops) and
    replayed 10 times; `frac` is algorithmic bytes / time against the same measured HBM peak (these kernels
    are issue- or MUFU-bound, the fraction says how far from the copy roofline that leaves them)."""
    from reslic_tcm_b200 import EntropyBottleneck
    from reslic_tcm_b200.stanh import GaussianConditionalStanh, compute_gap

    B, C, h, w = 256, 64, 16, 16
    g = torch.Generator(device=dev).manual_seed(1)
    sets = []
    for _ in range(3):
        mu = torch.randn(B, C, h, w, device=dev, generator=g)
        sigma = torch.exp(torch.empty(B, C, h, w, device=dev).uniform_(-3.0, 4.16, generator=g))
        y = mu + sigma * torch.randn(B, C, h, w, device=dev, generator=g)
        sets.append((y, sigma, mu, torch.randn(B, C, h, w, device=dev, generator=g), torch.randn(B, C, h, w, device=dev, generator=g)))
    n_y = B * C * h * w
    out = {}

    def timeit(name, fn, n, bpe, launches=12, reps=10):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(launches):
                fn(i)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (launches * reps)
        out[name] = {"us_per_launch": round(us, 2), "gelem_per_s": round(n / us * 1e-3, 1), "bytes_per_elem": bpe,
                     "frac": round(n * bpe / us * 1e-3 / peak, 3)}

    timeit("gc_bwd_noise", lambda i: ops.gc_backward(sets[i % 3][0], sets[i % 3][1], sets[i % 3][2], training=True,
                                                     g_yhat=sets[i % 3][3], g_lik=sets[i % 3][4], seed=3, offset=i % 3), n_y, 32)
    cfg = {"beta": 10.0, "num_sigmoids": 0, "extrema": 80, "symmetry": False, "trainable": False, "removing_mean": True}
    m = GaussianConditionalStanh(None, channels=C, gaussian_configuration=cfg).to(dev)
    m.stanh.update_state(dev)
    timeit("stanh_fwd_soft_beta10", lambda i: m.forward_fused(sets[i % 3][0], sets[i % 3][1], training=True, means=sets[i % 3][2],
                                                             want=("yhat", "lik")), n_y, 20)
    timeit("stanh_fwd_hard", lambda i: m.forward_fused(sets[i % 3][0], sets[i % 3][1], training=False, means=sets[i % 3][2],
                                                      want=("yhat", "lik")), n_y, 20)
    timeit("stanh_bwd_soft_beta10", lambda i: m._stanh_backward(sets[i % 3][0], sets[i % 3][1], sets[i % 3][2], True,
                                                               sets[i % 3][3], sets[i % 3][4]), n_y, 32)
    timeit("stanh_compute_gap", lambda i: compute_gap(m.stanh, sets[i % 3][0]), n_y, 4)
    Cz, hz = 192, 4
    mod = EntropyBottleneck(Cz).to(dev).train()
    synthetic.load_eb_parameters(mod, synthetic.eb_parameters())
    mm, bb, ff = mod._params()
    med = mod._medians_flat()
    zs = [torch.randn(B, Cz, hz, hz, device=dev, generator=g) * 4 for _ in range(3)]
    gs = [torch.randn(B, Cz, hz, hz, device=dev, generator=g) for _ in range(3)]
    n_z = B * Cz * hz * hz
    timeit("eb_fwd_noise", lambda i: ops.eb_forward(zs[i % 3], mm, bb, ff, med, training=True, want=("zhat", "lik"), seed=1, offset=i), n_z, 12)
    timeit("eb_bwd_noise", lambda i: ops.eb_backward(zs[i % 3], mm, bb, ff, med, training=True, g_zhat=gs[i % 3], g_lik=gs[(i + 1) % 3],
                                                    seed=1, offset=i), n_z, 20)
    return {"shape": {"y_slice": [B, C, h, w], "z": [B, Cz, hz, hz]}, "kernels": out}


def run_e2e(args, c, s, host, dev, world, B, elems_rank, kw, packed_slots=False):
    """Same step through the public API with HOST buffers (reslic_tcm_b200.pipeline.HostPipeline):
    every step copies y, mu, sigma, z from pinned host memory, runs the pass, and reads symbols and
    indexes (the rANS coder's input, tcm.py:551-552) plus the per-image bits back to pinned host
    memory.  Chunked over images so that H2D, kernels and D2H overlap."""
    import torch.distributed as dist

    from reslic_tcm_b200.pipeline import HostPipeline

    hp = HostPipeline(s["path"], B, c.y_hw, c.z_hw, with_indexes=c.with_indexes, training=c.training,
                      chunks=args.e2e_chunks, device=dev, num_pixels=c.num_pixels_per_image, seed=kw.get("seed", 0),
                      packed_slots=packed_slots)
    steps = max(3, min(args.steps, 60))
    for _ in range(3):
        out = hp.run(host)
    torch.cuda.synchronize()
    out = {k: v.clone() for k, v in out.items() if k != "done"}
    # correctness of the host round trip: same bits as the device-resident graph
    s["graph"].replay()
    torch.cuda.synchronize()
    ref_bits = s["res"]["bits"].double().cpu()
    # (chunked launches group the fp32 partials differently; chunked NOISE launches also draw different noise —
    # the Philox counter is the element index inside a launch — so that case is not comparable)
    comparable = not (c.training and len(hp.ranges) > 1)
    if comparable and not torch.allclose(out["bits"], ref_bits, rtol=1e-6):
        raise RuntimeError("e2e pipeline disagrees with the device-resident pass")
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        hp.run(host)                       # no synchronisation between batches: they stream
    torch.cuda.current_stream(dev).wait_stream(hp.s_d2h)     # the last batch's outputs are on the host
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": elems_rank * world / (ms / steps * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": hp.h2d_bytes,
            "d2h_bytes_per_step": hp.d2h_bytes, "ms_per_step": ms / steps, "steps": steps,
            "what": f"pinned host y/mu/sigma/z -> H2D -> 1+5{'+1' if packed_slots else ''} launches -> D2H {'/'.join(hp.out_names)}, "
                    f"{len(hp.ranges)} image chunks pipelined on 3 streams, batches double-buffered"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
