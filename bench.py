#!/usr/bin/env python
"""bench.py — entropy-model hot path throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 3] [--scaling strong|weak] [--impl ours|reference]

A "step" is one pass of the hot path over one GLOBAL batch of synthetic latents of the named config.  Default
workload: BASELINE.json configs[2] = TCM N=128, Kodak-shaped 768x512, batch 64, likelihood + bpp — the config the
metric names "at 1/2/4/8 GPUs".  Per step and rank: 1 EntropyBottleneck launch on z + 5 Gaussian-conditional
launches (one per channel slice, each waiting for its predecessor, as TCM drives them) + the per-image rate sum.
Prints ONE JSON line (rank 0).

* scaling    — "strong" (default): the config's batch is cut across the ranks (dist.shard_range: 64 -> 8 per GPU at
               N = 8; the north star's T1 / (N * TN)); "weak": every rank runs a full config-shaped batch.
* value      — latent elements (y + z) of the global batch / second, inputs resident in HBM, steps replayed as
               CUDA graphs, CUDA-event timed, max over ranks.  `--chains` independent batches are in flight at
               once (graph branches over different buffer sets; every batch is still a dependent launch chain).
* e2e        — same metric through the public API with HOST (pinned) buffers: H2D of y/mu/sigma/z and D2H of the
               step's results inside the timed region.
* roofline   — the Gaussian-conditional kernel alone: algorithmic bytes / its average launch duration (one
               dependent chain of slice launches replayed back to back); `per_config` repeats it for the other
               BASELINE configs (N = 1), `legs` carries config 5 sharded the same way (N > 1).
* cpu_baseline / --impl reference — the oracle restatement of the reference's own CPU op chain
               (oracle/compressai_ref.py) on the host cores, bounded sample.
Multi-GPU: the only exchange is the packed scalar rate row of each step.
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
DEFAULT_CONFIG = 3


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", type=int, default=DEFAULT_CONFIG, choices=sorted(synthetic.CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nbuf", type=int, default=0, help="rotating buffer sets (L2-cold inputs); 0 = as many as chains")
    ap.add_argument("--chains", type=int, default=0,
                    help="independent batches in flight at once (graph branches; buffer set s always runs on chain s %% chains); "
                         "0 = by launch size: 4, or 6 when a rank's slice launch is at most two waves of tiles (an 8-GPU shard)")
    ap.add_argument("--steps-per-graph", type=int, default=48, help="consecutive steps captured in one CUDA graph")
    ap.add_argument("--cpu-images", type=int, default=0, help="images in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--legs", default="auto", help="extra configs measured beside the main one: 'auto' (N=1: 2,4,5; N>1: 5), 'none', or a list '2,5'")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "fused", "branch", "nccl"],
                    help="rate exchange at N > 1 over peer memory (auto = peer = fused: the step's collecting launch publishes itself; "
                         "branch: a one-CTA publisher on a graph branch behind it — measured equal within 0.2 us per step) or nccl "
                         "(packed all-reduce, the fallback)")
    ap.add_argument("--shard-of", type=int, default=0,
                    help="development aid: run rank 0's shard of a G-rank strong-scaling job on ONE GPU (no exchange)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-whole-y", action="store_true")
    ap.add_argument("--no-training-kernels", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=1,
                    help="image chunks per batch in the host pipeline (batches are double-buffered, so 1 streams best)")
    return ap.parse_args(argv)


def bytes_per_y_elem(c: synthetic.Config) -> int:
    """Algorithmic HBM bytes per y element of one GC launch: read y, mu,
    sigma (12) + write y_hat, L (8) [+ symbols, indexes (8)] [+ noisy y (4) in training]."""
    return 12 + 8 + (8 if c.with_indexes else 0) + (4 if c.training else 0)


def workload_config(c: synthetic.Config, world: int, scaling: str) -> dict:
    """The `config` object of the JSON line — identical for both arms (the reference arm steps over a bounded
    sample of the same workload and says so in `cpu_baseline.sample`)."""
    return {"workload": c.name, "cfg": c.cfg, "global_batch": c.batch * (world if scaling == "weak" else 1),
            "y_shape_per_image": [synthetic.M_LATENT, *c.y_hw], "z_shape_per_image": [synthetic.Z_CHANNELS, *c.z_hw],
            "mode": "training (noise)" if c.training else ("eval round + symbols + indexes" if c.with_indexes else "eval round, likelihood + bpp")}


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


def gpu_reference_throughput(c: synthetic.Config, n_images: int, dev, reps: int = 5):
    """The same unfused op chain (oracle/compressai_ref.py tcm_entropy_step: the reference's op order, stock PyTorch
    eager kernels) run ON THE B200 — the "reference-on-GPU" line of SURVEY.md section 8(d), the fairer denominator for what
    the fused kernels buy on the same silicon.  Inputs resident, CUDA events, best of `reps` after one warm-up."""
    from oracle import compressai_ref as cr

    batch = {k: v.to(dev) for k, v in synthetic.make_batch(c.cfg, range(n_images), with_noise=c.training).items()}
    eb = ref_eb({k: v.to(dev) for k, v in synthetic.eb_parameters().items()})
    table = synthetic.scale_table(dev)
    elems = n_images * (c.y_elems_per_image + c.z_elems_per_image)

    def once():
        with torch.no_grad():
            cr.tcm_entropy_step(batch["y"], batch["mu"], batch["sigma"], batch["z"], eb, table, training=c.training,
                                with_indexes=c.with_indexes, num_pixels=n_images * c.num_pixels_per_image,
                                noise_y=batch.get("noise_y"), noise_z=batch.get("noise_z"))

    once()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        once()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return elems / (best * 1e-3) / 1e6, best, elems


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
    sample = (f"each step = {n_img} of the {c.batch} images of config {c.cfg} ({elems} latent elements; throughput is per element, "
              f"so the sample size does not enter the ratio), {steps} steps")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(c, max(world, args.gpus, 1), args.scaling),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "cpu": cpu_model()},
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


# --------------------------------------------------------------------------- host placement
def _parse_cpulist(txt: str) -> set:
    cpus = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(dev) -> tuple:
    """(cpu set, source) of the CPUs local to `dev`: NVML's ideal affinity first (what `nvidia-smi topo -m`
    prints — it knows the real topology even where the container's sysfs reports node 0 for every GPU), the
    PCI device's sysfs local_cpulist second."""
    p = torch.cuda.get_device_properties(dev)
    try:
        import pynvml as nv

        nv.nvmlInit()
        try:
            h = nv.nvmlDeviceGetHandleByUUID("GPU-" + str(p.uuid))
        except Exception:
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(p.uuid)).encode())
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        if cpus:
            return cpus, "nvml"
    except Exception:
        pass
    bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    return _parse_cpulist(open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read()), "sysfs"


def bind_to_gpu_numa(dev) -> dict:
    """Multi-rank runs: pin this process to the CPUs local to its GPU so that its pinned host buffers are
    first-touched on that NUMA node and the e2e copies do not cross sockets.  Returns what was done."""
    info = {"bound": False}
    try:
        cpus, src = gpu_local_cpus(dev)
        allowed = os.sched_getaffinity(0)
        info.update(source=src, local_cpus=len(cpus), allowed_cpus=len(allowed))
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
        info["cpus_used"] = len(use or allowed)
        nodes = set()
        for n in os.listdir("/sys/devices/system/node"):
            if n.startswith("node") and _parse_cpulist(open(f"/sys/devices/system/node/{n}/cpulist").read()) & (use or allowed):
                nodes.add(int(n[4:]))
        info["numa_nodes"] = sorted(nodes)
    except Exception as e:      # topology not exposed in this container: run unbound
        info["error"] = repr(e)
    print(f"[bench] host placement: {info}", file=sys.stderr)
    return info


# --------------------------------------------------------------------------- GPU arm
class Workload:
    """One rank's share of one BASELINE config on the device: `nbuf` rotating buffer sets (each its own
    TcmEntropyPath = static outputs + rate workspace), graphs of consecutive steps, the kernel-only leg."""

    def __init__(self, c: synthetic.Config, images: range, dev, nbuf: int, params, pin_host: bool = False, path_factory=None):
        from reslic_tcm_b200.pipeline import TcmEntropyPath

        self.c, self.dev, self.B = c, dev, len(images)
        self.host = synthetic.make_batch(c.cfg, images, with_noise=False, pin=pin_host)
        self.kw = dict(training=c.training, with_indexes=c.with_indexes, num_pixels=c.num_pixels_per_image, seed=1234)
        self.sets = []
        for _ in range(max(1, nbuf)):
            if path_factory is not None:          # another model's entropy pass over the same latents (stanh_step_leg)
                path = path_factory().to(dev).eval()
                synthetic.load_eb_parameters(path.entropy_bottleneck, params)
            else:
                path = TcmEntropyPath().to(dev).eval()
                synthetic.load_eb_parameters(path.entropy_bottleneck, params)
                path.gaussian_conditional.scale_table = synthetic.scale_table(dev)
            inp = {k: self.host[k].to(dev, non_blocking=True) for k in ("y", "mu", "sigma", "z")}
            torch.cuda.synchronize()
            res = path.forward(inp["y"], inp["mu"], inp["sigma"], inp["z"], **self.kw)   # warm-up: lazy init, LUT, allocator
            self.sets.append({"path": path, "inp": inp, "res": res})
        torch.cuda.synchronize()
        self.y_elems, self.z_elems = self.B * c.y_elems_per_image, self.B * c.z_elems_per_image
        self.bpe = bytes_per_y_elem(c)
        self._side = []

    def set_bytes(self) -> int:
        return self.bpe * self.y_elems + 12 * self.z_elems

    def step(self, i: int, **over):
        s = self.sets[i % len(self.sets)]
        kw = dict(self.kw, offset=8 * i, **over)          # noise mode: every captured step draws its own Philox field
        return s["path"].forward(s["inp"]["y"], s["inp"]["mu"], s["inp"]["sigma"], s["inp"]["z"], **kw)

    def capture(self, n_steps: int, chains: int = 1, first: int = 0, adapter=None, **over):
        """ONE graph of steps first .. first+n_steps-1.  Consecutive launches of a step are PDL edges; with
        chains > 1 the steps are dealt onto that many forked capture streams — buffer set s always on chain
        s % chains, so two steps over the same buffers stay ordered — i.e. `chains` independent batches are in
        flight at once, each still a dependent chain.  `adapter` (an exchange of bench.py) adds the multi-GPU rate
        exchange: per-step keyword arguments, work captured behind a step, a prologue branch, streams to join and an
        epilogue behind the join."""
        chains = max(1, min(chains, len(self.sets)))
        while len(self._side) < chains - 1:
            self._side.append(torch.cuda.Stream(device=self.dev))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream(self.dev)
            lanes = [cur] + self._side[:chains - 1]
            for s in lanes[1:]:
                s.wait_stream(cur)
            extra = adapter.prologue(self, cur, n_steps) if adapter is not None else []
            for j in range(first, first + n_steps):
                si = j % len(self.sets)
                with torch.cuda.stream(lanes[si % chains]):
                    kw = dict(over, **adapter.step_kwargs(j - first)) if adapter is not None else over
                    self.step(j, **kw)
                    if adapter is not None:
                        adapter.after_step(j - first, self.sets[si], si % chains)
            for s in lanes[1:] + list(extra) + (adapter.join_streams() if adapter is not None else []):
                cur.wait_stream(s)
            if adapter is not None:
                adapter.epilogue(n_steps)      # behind the join
        return g

    def drain_deferred(self):
        for s in self.sets:
            b = s["path"].buffers(s["inp"]["y"], s["inp"]["z"], self.c.with_indexes, self.c.training)
            ops.rate_finalize(b["workspace"], self.B, bits=b["bits"])

    def gc_only_us(self, fuse: bool, chains: int = 1, min_reps: int = 50):
        """Only the GC launches of a step (5 per-slice, or 1 over the whole y) over the rotating buffer sets, captured as
        one graph of >= 60 launches so that the launch duration is not diluted by graph-replay boundaries; `chains` as in
        the timed region (1 = ONE dependent chain, the kernel in isolation).  The rate stays deferred in the workspace and
        is drained after the timed region.  Returns (us per launch = elapsed / launches, launches per step)."""
        n_launch = 1 if fuse else synthetic.NUM_SLICES
        over = dict(skip_z=True, fuse_slices=fuse, defer_rate=True)
        n_steps = max(len(self.sets), -(-60 // n_launch))
        n_steps += (-n_steps) % len(self.sets)
        for i in range(len(self.sets)):
            self.step(i, **over)
        self.drain_deferred()
        torch.cuda.synchronize()
        g = self.capture(n_steps, chains=chains, **over)
        per_graph = n_steps * n_launch
        reps = max(3, -(-min_reps * n_launch // per_graph))
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(reps):
            g.replay()
        r1.record()
        self.drain_deferred()            # outside the timed region: this leg times the kernel alone
        torch.cuda.synchronize()
        return r0.elapsed_time(r1) * 1e3 / (reps * per_graph), n_launch


def balanced_group(steps: int, max_group: int) -> int:
    """Steps per CUDA graph: the timed region is cut into equal graphs of at most `max_group` steps (200 -> 5 x 40,
    20 -> 1 x 20), so that no short tail graph runs with most of its branches empty."""
    n_graphs = max(1, -(-steps // max(1, max_group)))
    return max(1, -(-steps // n_graphs))


def pick_chains(c: synthetic.Config, images: int, group: int) -> int:
    """Independent batches in flight.  Launches of at most two waves of tiles (740 CTA slots of 1024 elements on B200: an
    8-GPU shard of a slice) are latency-bound and gain from more batches in flight (measured, 8 images of config 3:
    4 -> 6 batches 16.1 -> 15.4 us per step); the full 64-image batch loses 1.4 % with 6 and keeps 4.  Among the
    candidates the one that deals the `group` steps of a graph evenly onto its branches wins (20 steps on 6 branches
    leave two branches a step longer than the rest: 18.2 instead of 15.8 us per step)."""
    slice_elems = images * c.y_elems_per_image // synthetic.NUM_SLICES
    cands = (6, 5, 4) if slice_elems <= 2 * 740 * 1024 else (4, 3)
    return max(cands, key=lambda k: (group / (k * -(-group // k)), k))


def load_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    return peak, src


def load_traffic():
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            db = json.load(open(os.path.join(ROOT, "profiles", name))).get("gc_fwd_kernel", [])
            return (db if isinstance(db, list) else [db]), name
        except Exception:
            continue
    return [], None


def roof_obj(w: Workload, us: float, n_launch: int, peak: float, peak_src: str, traffic_db, what: str) -> dict:
    elems = w.y_elems // n_launch
    achieved = w.bpe * elems / (us * 1e-6) / 1e9
    traffic = None      # dram read+write bytes per launch from the committed ncu --set full captures
    for t in traffic_db:
        if t.get("config") == w.c.cfg and t.get("elems_per_launch") == elems:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": "gc_fwd_kernel", "bytes_per_elem": w.bpe, "elems_per_launch": elems,
            "us_per_launch": us, "peak_source": peak_src, "gc_melem_per_s": elems / us, "launch": what}


SLICE_WHAT = "one 64-channel slice of y per launch, each launch waiting for its predecessor (TCM's call pattern, tcm.py:443-457)"
WHOLE_WHAT = "all 320 channels of y in one launch (models whose mu/sigma exist for every channel at once)"


def roof_pair(w: Workload, fuse: bool, chains: int, peak, peak_src, traffic_db, what: str) -> dict:
    """The roofline object of the GC kernel as the timed region runs it (`chains` independent batches in flight:
    us_per_launch = elapsed / launches, i.e. the HBM rate the path sustains) with the kernel in isolation — ONE dependent
    chain, nothing else on the GPU — beside it as `single_chain`."""
    chains = max(1, min(chains, len(w.sets)))
    us1, n = w.gc_only_us(fuse, 1)
    single = roof_obj(w, us1, n, peak, peak_src, traffic_db, what + "; one dependent chain alone on the GPU")
    if chains == 1:
        return single
    usc, n = w.gc_only_us(fuse, chains)
    r = roof_obj(w, usc, n, peak, peak_src, traffic_db,
                 what + f"; {chains} independent batches in flight as in the timed region (us_per_launch = elapsed / launches)")
    r["single_chain"] = {k: single[k] for k in ("us_per_launch", "achieved", "frac", "gc_melem_per_s")}
    return r


class Timer:
    """Timed loop over graphs of `group` steps (+ one partial graph for the tail), max over ranks."""

    def __init__(self, w: Workload, group: int, chains: int, world: int, exchange=None, **over):
        self.w, self.group, self.chains, self.world, self.ex = w, max(1, group), chains, world, exchange
        self.over = dict(over)
        self.graphs = {}

    def graph(self, n: int):
        if n not in self.graphs:
            self.graphs[n] = self.w.capture(n, self.chains, adapter=self.ex, **self.over)
        return self.graphs[n]

    def run(self, n_steps: int):
        """Enqueue n_steps steps.  Every graph starts at buffer set 0, so a partial graph is just a shorter one."""
        left = n_steps
        while left > 0:
            n = min(left, self.group)
            self.last_n = n
            if self.ex is not None:
                self.ex.before_graph(n)
            self.graph(n).replay()
            if self.ex is not None:
                self.ex.after_graph(n)
            left -= n

    def timed(self, steps: int, warmup: int, barrier, sampler=None):
        import torch.distributed as dist

        # warm-up: at least `warmup` steps, and every graph the timed region will replay at least once (the first launch of
        # a CUDA graph uploads it: a graph first replayed inside the timed region costs that region hundreds of microseconds)
        sizes = [n for n in dict.fromkeys([min(steps, self.group), steps % self.group if steps > self.group else 0]) if n]
        done = 0
        for n in sizes:
            self.run(n)
            done += n
        if done < max(warmup, 3):
            self.run(max(warmup, 3) - done)
            done = max(warmup, 3)
        self.warmup_steps = done
        barrier()
        if self.ex is not None:
            self.ex.verify()          # a broken exchange fails here, after the warm-up, not after minutes of timed-out reads
            barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.active.set()
        # The region is timed ON THE DEVICE: a ~1 ms spin kernel goes first so that the start event and the first graph
        # are already queued when the GPU reaches them — otherwise the host's cudaGraphLaunch of the first graph (tens of
        # microseconds for ~140 nodes, more with eight ranks launching at once; the slowest rank sets the time) sits
        # between the event and the first kernel, which is host latency, not step time (it is >= 5 % of the driver's
        # 20-step region on an 8-GPU shard).  The spin is before the start event: not timed.
        torch.cuda._sleep(2_000_000)
        e0.record()
        self.run(steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.w.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if sampler is not None:
            # the timed region may be shorter than an NVML sample: keep the identical loop running (untimed) until the
            # sampler has seen >= 1 s of it — the same number of graphs on every rank (ms is the max over ranks), so the
            # exchange stays in step
            extra = int(max(0.0, 1000.0 - ms) / max(ms / steps * self.group, 1e-6)) + 1
            for k in range(extra):
                self.run(self.group)
                if k % 64 == 63:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            sampler.active.clear()
        return ms / steps


class NcclExchange:
    """Rate exchange by NCCL (the checked fallback): the steps of one graph share ONE packed all-reduce of a
    [group, 4] float64 matrix (bits, sq_err, pixels, images), packed by tiny kernels inside the graph."""

    name = "nccl all-reduce of one packed [steps, 4] float64 matrix per graph"

    def __init__(self, w: Workload, group: int):
        from reslic_tcm_b200 import dist as rdist

        self.red = rdist.RateReducer(w.dev, slots=group)
        self.red.set_static(0.0, w.B * w.c.num_pixels_per_image, w.B)
        self.work = None

    def prologue(self, w, cur, n):
        return []

    def step_kwargs(self, j):
        return {}

    def after_step(self, j, s, chain):
        self.red.pack_bits(s["res"]["bits"], slot=j)

    def join_streams(self):
        return []

    def epilogue(self, n):
        pass

    def before_graph(self, n):
        if self.work is not None:         # stream-level wait: the previous collective has read the matrix
            self.work.wait()
            self.work = None

    def after_graph(self, n):
        self.work = self.red.all_reduce(async_op=True)

    def verify(self):
        pass

    def result(self):
        if self.work is not None:
            self.work.wait()
        torch.cuda.synchronize()
        return self.red.result(0)

    def result_set(self, w, last_n):
        return 0                                   # result() is step 0 of the last graph = buffer set 0

    def close(self):
        pass


class PeerExchange:
    """Rate exchange over peer memory (reslic_tcm_b200.dist.PeerRateExchange): every step stores its packed row into every
    rank's buffer over NVLink; no collective kernel exists.  Two forms of the writer:
      fused   — the step's collecting launch publishes (reslic_gc_desc.exchange): nothing but that launch runs;
      branch  — the collecting launch stays the plain kernel and a one-CTA publisher (reslic_rate_exchange_publish_f64)
                sits on a graph branch of its own behind it, so the chain's next step does not wait for the publish.
    Every graph opens, on another branch, with one tiny read kernel over the rows the PREVIOUS graph published — off the
    steps' critical path, and what keeps a rank from running a ring ahead."""

    def __init__(self, w: Workload, group: int, form: str):
        from reslic_tcm_b200 import dist as rdist

        self.form = form
        self.name = ("packed rate row stored to every rank over NVLink, no collective kernel; writer: " +
                     ("the step's collecting launch itself (fused)" if form == "fused" else
                      "a one-CTA publisher on a graph branch behind the collecting launch") +
                     "; one read kernel per graph on a side branch, one graph behind")
        self.ex = rdist.PeerRateExchange(w.dev, ring=max(256, 8 * group))
        self.ex.set_static(w.B * w.c.num_pixels_per_image, w.B)
        self.published = 0
        self.rd = torch.cuda.Stream(device=w.dev)
        self.pub = {}
        self.behind = torch.zeros(max(group, 64), 4, dtype=torch.float64, device=w.dev)

    def prologue(self, w, cur, n):
        self.rd.wait_stream(cur)
        with torch.cuda.stream(self.rd):
            self.ex.read_behind(n, out=self.behind[:n])
        self._used = set()
        return [self.rd]

    def step_kwargs(self, j):
        # the batch's number inside this graph: its exchange slot does not depend on which chain finishes first
        return {"exchange": self.ex, "exchange_step": j, "exchange_advance": False} if self.form == "fused" else {}

    def after_step(self, j, s, chain):
        if self.form == "fused":
            return
        lane = torch.cuda.current_stream(self.ex.device)
        pub = self.pub.setdefault(chain, torch.cuda.Stream(device=self.ex.device))
        pub.wait_stream(lane)                     # behind this step's collecting launch only
        with torch.cuda.stream(pub):
            self.ex.publish(s["res"]["bits"], step=j)
        self._used.add(chain)

    def join_streams(self):
        return [self.pub[c] for c in sorted(self._used)]

    def epilogue(self, n):
        self.ex.advance(n)                        # every replay publishes the next n steps

    def before_graph(self, n):
        pass

    def after_graph(self, n):
        self.published += n

    def verify(self):
        torch.cuda.synchronize()
        self.ex.check()

    def result(self):
        self.ex.read_step = max(self.published - 1, 0)        # the last step, by its absolute number
        row = self.ex.read(1)[0]
        self.verify()
        return self.ex.result(row.tolist())

    def result_set(self, w, last_n):
        # result() is the LAST step of the last graph; in noise mode every step of a graph draws its own Philox field,
        # so the local row to compare with must come from the buffer set that step wrote
        return (last_n - 1) % len(w.sets)

    def close(self):
        self.ex.close()


def measure_config(c, images, dev, args, params, world, global_elems, peak, peak_src, traffic_db, barrier,
                   sampler=None, exchange_factory=None, steps=None, pin_host=False, light=False, sim_world=1):
    """value / ms_per_step / roofline (+ whole-y) of one config for this rank's `images`."""
    steps = steps or args.steps
    group = balanced_group(steps, args.steps_per_graph)
    chains = args.chains or pick_chains(c, len(images), group)
    nbuf = args.nbuf or chains
    w = Workload(c, images, dev, nbuf, params, pin_host=pin_host)
    ex = exchange_factory(w) if exchange_factory is not None else None
    tm = Timer(w, group, chains, world, ex)
    ms_step = tm.timed(steps, args.warmup, barrier, sampler)
    out = {"workload": c.name, "cfg": c.cfg, "images_per_gpu": w.B, "value": global_elems / (ms_step * 1e-3) / 1e6, "unit": UNIT,
           "ms_per_step": ms_step, "steps": steps, "warmup_steps": tm.warmup_steps, "launches_per_step": 1 + synthetic.NUM_SLICES,
           "batches_in_flight": min(chains, nbuf), "buffer_sets": nbuf, "steps_per_graph": group}
    if ex is not None:
        # the exchanged global row against the checked fallback: an NCCL all-reduce of the same step's local row
        import torch.distributed as dist

        got = ex.result()
        local = torch.tensor([float(w.sets[ex.result_set(w, tm.last_n)]["res"]["bits"].double().sum()), 0.0,
                              float(w.B * c.num_pixels_per_image), float(w.B)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(local)
        want = local.tolist()
        out["exchange_check"] = {"exchange": got, "nccl_all_reduce": {"bits": want[0], "pixels": want[2], "images": want[3]},
                                 "match": bool(abs(got["bits"] - want[0]) <= 1e-9 * abs(want[0]) and got["pixels"] == want[2] and got["images"] == want[3])}
        out["exchange_name"] = ex.name
    if ex is not None or sim_world > 1:
        # where the step's time goes on this rank: the GC launches alone in the same launch pattern (kernel), the same step
        # without the exchange (z launch, launch gaps, graph ramp on top of the kernel), and what the exchange adds
        t_noex = Timer(w, group, chains, world, None).timed(steps, args.warmup, barrier) if ex is not None else ms_step
        us_k, n_k = w.gc_only_us(False, chains)
        out["step_budget"] = {
            "us_per_step": ms_step * 1e3, "gc_kernels_us": us_k * n_k, "bottleneck_launch_gaps_ramp_us": t_noex * 1e3 - us_k * n_k,
            "exchange_us": (ms_step - t_noex) * 1e3 if ex is not None else None,
            "what": "per rank, max over ranks of each timed loop; gc_kernels = 5 x elapsed / launches of the GC launches "
                                      "alone with the same batches in flight; exchange = step with the fused publish + read - the same step without"}
    if ex is not None:
        ex.close()
    out["roofline"] = roof_pair(w, False, chains, peak, peak_src, traffic_db, SLICE_WHAT)
    if not args.no_whole_y:
        tw = Timer(w, group, chains, world, None, fuse_slices=True)
        wms = tw.timed(min(steps, 100) if light else steps, min(args.warmup, 6), barrier)
        out["whole_y"] = {"value": global_elems / (wms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": wms, "launches_per_step": 2,
                          "roofline": roof_pair(w, True, chains, peak, peak_src, traffic_db, WHOLE_WHAT)}
    return w, out


def run_ours(args):
    import torch.distributed as dist

    from reslic_tcm_b200 import _cabi, dist as rdist

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
    placement = bind_to_gpu_numa(dev) if world > 1 else None     # before any pinned allocation (first touch)
    params = synthetic.eb_parameters()
    peak, peak_src = load_peak()
    traffic_db, traffic_src = load_traffic()
    sim = args.shard_of if (args.shard_of > 1 and world == 1) else 0
    eff_world = sim or world

    def shard(c):
        """(this rank's images, latent elements of the global batch)."""
        per_image = c.y_elems_per_image + c.z_elems_per_image
        if args.scaling == "weak":
            return range(rank * c.batch, (rank + 1) * c.batch), c.batch * per_image * eff_world
        if c.batch < eff_world:
            raise SystemExit(f"config {c.cfg} has {c.batch} image(s): it does not shard over {eff_world} ranks (replicas only)")
        return rdist.shard_range(c.batch, rank, eff_world), c.batch * per_image

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

    exchange_name, exchange_factory = None, None
    if world > 1 or args.exchange in ("peer", "fused", "branch"):   # (on one GPU: the whole exchange path with world = 1)
        def exchange_factory(w):
            return make_exchange(args, w, rank, world)

    # ---- main leg
    c = synthetic.CONFIGS[args.config]
    images, global_elems = shard(c)
    w, main = measure_config(c, images, dev, args, params, world, global_elems, peak, peak_src, traffic_db, barrier,
                             sampler=sampler, exchange_factory=exchange_factory, pin_host=not args.no_e2e, sim_world=eff_world)
    exchange_name = main.pop("exchange_name", None)

    # ---- e2e leg: public API with host buffers
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, c, w, dev, world, global_elems, packed_slots=c.with_indexes)
    bits0 = w.sets[0]["res"]["bits"].double().cpu()

    # ---- training-path kernels on a config-5 slice (single-GPU runs): backward, noise-mode bottleneck, STanH
    training_kernels = None
    if world == 1 and not sim and not args.no_training_kernels:
        training_kernels = training_kernels_leg(dev, peak)
    del w
    torch.cuda.empty_cache()
    # ---- config 5 as the reference's STanH model runs it (annealed soft quantization), single-GPU runs
    stanh_step = None
    if not args.no_training_kernels and synthetic.CONFIGS[5].batch >= eff_world:
        imgs5, gel5 = shard(synthetic.CONFIGS[5])
        stanh_step = stanh_step_leg(args, dev, params, peak, imgs5, gel5, world, barrier)
        torch.cuda.empty_cache()

    # ---- the other BASELINE configs: per-config roofline objects (N = 1) / config 5 sharded the same way (N > 1)
    if args.legs == "auto":
        leg_cfgs = [k for k in ((2, 4, 5) if eff_world == 1 else (5,)) if k != c.cfg]
    elif args.legs == "none":
        leg_cfgs = []
    else:
        leg_cfgs = [int(k) for k in args.legs.split(",") if k.strip() and int(k) != c.cfg]
    legs = {}
    for k in leg_cfgs:
        ck = synthetic.CONFIGS[k]
        imgs, gel = shard(ck)
        wk, legs[str(k)] = measure_config(ck, imgs, dev, args, params, world, gel, peak, peak_src, traffic_db, barrier,
                                          exchange_factory=exchange_factory, steps=min(args.steps, 100), light=True,
                                          pin_host=(k == 2 and not args.no_e2e))
        legs[str(k)].pop("exchange_name", None)
        if k == 2 and not args.no_e2e:      # the compress-path config: host round trip with the packed rANS slots
            legs[str(k)]["e2e"] = run_e2e(args, ck, wk, dev, world, gel, packed_slots=True)
        del wk
        torch.cuda.empty_cache()

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
        try:
            n_gpu = min(c.batch, 16)
            gv, gms, gel = gpu_reference_throughput(c, n_gpu, dev)
            cpu["on_gpu"] = {"value": gv, "unit": UNIT, "ms": gms,
                             "what": f"the same unfused op chain (oracle port, reference op order) as stock PyTorch {torch.__version__} eager CUDA "
                                     f"kernels on this B200, {n_gpu} of {c.batch} images ({gel} latent elements) resident, best of 5"}
        except Exception as e:      # a baseline line, never the product: report and go on
            cpu["on_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()

    if rank == 0:
        cfg = workload_config(c, world, args.scaling)
        cfg.update({
            "images_per_gpu": main["images_per_gpu"], "launches_per_step": main["launches_per_step"],
            "step": f"1 EB + 5 GC launches per step (each GC launch waits for its predecessor), {main['steps_per_graph']} steps per CUDA graph, "
                    f"{main['batches_in_flight']} independent batches in flight (graph branches over different buffer sets)",
            "l2": f"{main['buffer_sets']} rotating buffer sets of {(bytes_per_y_elem(c) * main['images_per_gpu'] * c.y_elems_per_image + 12 * main['images_per_gpu'] * c.z_elems_per_image) / 1e6:.0f} MB each "
                  f"per GPU; inputs + outputs of one step exceed the 126 MB L2 only while a rank holds >= 13 images of this config — "
                  f"smaller shards are L2-assisted, which is what a real 8-GPU run sees too",
            "bpp_mean_rank0": float(bits0.mean()) / c.num_pixels_per_image,
        })
        if exchange_name:
            cfg["exchange"] = exchange_name
        if sim:
            cfg["simulated_shard_of"] = sim
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "warmup_steps_run": main["warmup_steps"],      # asked for / actually run (every graph of the timed region is replayed once first)
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": main["roofline"], "whole_y": main.get("whole_y"),
            "per_config": {str(c.cfg): {k: main[k] for k in ("workload", "value", "ms_per_step", "roofline", "whole_y") if k in main}, **legs},
            "legs": {k: {"workload": v["workload"], "value": v["value"], "ms_per_step": v["ms_per_step"], "images_per_gpu": v["images_per_gpu"],
                         "scaling": args.scaling} for k, v in legs.items()},
            "training_kernels": training_kernels, "stanh_step": stanh_step, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": main["launches_per_step"] * args.steps, "clocks": clocks, "traffic_source": traffic_src,
        }
        if placement is not None:
            line["host_placement"] = placement
        if main.get("exchange_check") is not None:
            line["exchange_check"] = main["exchange_check"]
        if main.get("step_budget") is not None:
            line["step_budget"] = main["step_budget"]
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def make_exchange(args, w, rank, world, form=None):
    """The rate exchange of an N > 1 run: peer stores fused into the collecting launch (`form` "branch": the one-CTA
    publisher behind it, for passes whose collecting launch is not a Gaussian-conditional kernel); the NCCL all-reduce is
    the fallback (--exchange nccl, or when the peer buffers cannot be mapped — then every rank must fall back together)."""
    import torch.distributed as dist

    form = form or ("branch" if args.exchange == "branch" else "fused")
    if world == 1:
        return PeerExchange(w, args.steps_per_graph, form)
    if args.exchange != "nccl":
        ok, ex = 1, None
        try:
            ex = PeerExchange(w, args.steps_per_graph, form)
        except Exception as e:      # CUDA IPC not permitted in this container, no peer access, ...
            if args.exchange != "auto":
                raise
            print(f"[bench] peer exchange unavailable ({e!r}); falling back to NCCL", file=sys.stderr)
            ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device=w.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 1:
            return ex
        if ex is not None:
            ex.close()
    return NcclExchange(w, args.steps_per_graph)


def training_kernels_leg(dev, peak):
    """Per-launch times of the training-path kernels on one config-5 slice (256 patches of 256x256: y slice
    [256, 64, 16, 16], z [256, 192, 4, 4]) — the Gaussian-conditional backward, the noise-mode bottleneck
    forward / backward and the STanH family (soft beta = 10, hard, backward, compute_gap).  Each op is
    captured as a graph of 12 launches over 3 rotating input sets (403 MB > L2 for the y-shaped ops), each launch
    writing its own output tensors (12 distinct sets, so stores reach HBM), and replayed 10 times; `frac` is algorithmic
    bytes / time against the same measured HBM peak (these kernels are issue- or MUFU-bound, the fraction says how far
    from the copy roofline that leaves them)."""
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
        keep = []          # every captured launch keeps its OWN outputs alive: the graph's allocator would otherwise hand
        with torch.cuda.graph(gr):      # all 12 launches the same block, whose stores then never leave L2
            for i in range(launches):
                keep.append(fn(i))
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


def stanh_step_leg(args, dev, params, peak, images, global_elems, world, barrier, beta: float = 10.0):
    """BASELINE config 5 ("noise + annealed soft quantization") as the reference's STanH model runs its entropy pass
    (src/models/stanh/tcm_stanh.py:396-451; reslic_tcm_b200.pipeline.TcmStanhEntropyPath): noise-mode bottleneck on z,
    five GaussianConditionalStanh launches (soft quantization at beta about the predicted mean, variable-bin
    likelihood, ste value, rate) and one pass over the whole y for quantize("training") + compute_gap — 7 launches per
    step over this rank's `images` (the batch cut over the ranks as in the main leg), rotating buffer sets (each step's
    reads + writes = 28 B per y element), steps dealt onto graph branches as in the main timed region, max over ranks.
    N > 1: the collecting launch is a STanH kernel, so the rate row is published by the one-CTA publisher on a graph
    branch behind it (PeerExchange "branch") and checked against an NCCL all-reduce.  N = 1: the same steps as ONE
    dependent chain beside it.  `frac` = algorithmic bytes (y, mu, sigma read; soft value, likelihood, ste value
    written; y read again by the gap pass) / time / the measured HBM peak: these kernels are issue-bound (DESIGN.md
    section 3.3), the fraction says how far from the copy roofline that leaves the STanH step."""
    from reslic_tcm_b200.pipeline import TcmStanhEntropyPath

    c = synthetic.CONFIGS[5]
    cfg = {"beta": beta, "num_sigmoids": 0, "extrema": 80, "symmetry": False, "trainable": False, "removing_mean": True}
    steps = min(args.steps, 96)
    group = balanced_group(steps, min(args.steps_per_graph, 24))
    chains = args.chains or pick_chains(c, len(images), group)
    nbuf = args.nbuf or chains
    w = Workload(c, images, dev, nbuf, params, path_factory=lambda: TcmStanhEntropyPath(cfg, channels=64))
    w.kw = dict(training=True, num_pixels=c.num_pixels_per_image, seed=1234)
    bytes_step = 28 * w.y_elems + 12 * w.z_elems
    out = {"workload": c.name + "_stanh_soft", "beta": beta, "launches_per_step": 7, "bytes_per_y_elem": 28,
           "images_per_gpu": w.B, "buffer_sets": nbuf, "steps": steps, "steps_per_graph": group, "scaling": args.scaling}
    ex = make_exchange(args, w, 0, world, form="branch") if world > 1 else None
    tm = Timer(w, group, chains, world, ex)
    ms = tm.timed(steps, min(args.warmup, 8), barrier)
    us = ms * 1e3
    out["in_flight"] = {"us_per_step": round(us, 2), "value": round(global_elems / us, 1), "unit": UNIT,
                        "achieved_gbs_per_gpu": round(bytes_step / us * 1e-3, 1), "frac": round(bytes_step / us * 1e-3 / peak, 3),
                        "batches_in_flight": min(chains, nbuf)}
    if ex is not None:
        import torch.distributed as dist

        got = ex.result()
        local = torch.tensor([float(w.sets[ex.result_set(w, tm.last_n)]["res"]["bits"].double().sum()),
                              float(w.B * c.num_pixels_per_image), float(w.B)], dtype=torch.float64, device=dev)
        dist.all_reduce(local)
        want = local.tolist()
        out["exchange_check"] = {"exchange": got, "nccl_all_reduce": {"bits": want[0], "pixels": want[1], "images": want[2]},
                                 "match": bool(abs(got["bits"] - want[0]) <= 1e-9 * abs(want[0]) and got["pixels"] == want[1]
                                               and got["images"] == want[2])}
        out["exchange"] = ex.name
        ex.close()
    else:
        us1 = Timer(w, group, 1, world, None).timed(steps, min(args.warmup, 8), barrier) * 1e3
        out["single_chain"] = {"us_per_step": round(us1, 2), "value": round(global_elems / us1, 1), "unit": UNIT,
                               "achieved_gbs_per_gpu": round(bytes_step / us1 * 1e-3, 1), "frac": round(bytes_step / us1 * 1e-3 / peak, 3),
                               "batches_in_flight": 1}
    res = w.sets[0]["res"]
    out["bpp_mean_rank"] = float((res["bits"].double() / c.num_pixels_per_image).mean())
    out["gap_rank"] = float(w.sets[0]["path"].gap(res, w.y_elems))
    return out


def run_e2e(args, c, w: Workload, dev, world, global_elems, packed_slots=False):
    """Same step through the public API with HOST buffers (reslic_tcm_b200.pipeline.HostPipeline):
    every step copies this rank's y, mu, sigma, z from pinned host memory, runs the pass, and reads the
    step's results back to pinned host memory — the per-image bits, and for the compress-path configs the
    rANS coder's input (tcm.py:551-552) as ONE packed (start, range) slot per symbol."""
    import torch.distributed as dist

    from reslic_tcm_b200.pipeline import HostPipeline

    s = w.sets[0]
    if packed_slots:
        s["path"].gaussian_conditional.update()       # CDF tables of the scale table in place (setup, untimed)
    hp = HostPipeline(s["path"], w.B, c.y_hw, c.z_hw, with_indexes=c.with_indexes, training=c.training,
                      chunks=args.e2e_chunks, device=dev, num_pixels=c.num_pixels_per_image, seed=w.kw.get("seed", 0),
                      packed_slots=packed_slots)
    host = w.host
    steps = max(3, min(args.steps, 40))
    for _ in range(3):
        out = hp.run(host)
    torch.cuda.synchronize()
    out = {k: v.clone() for k, v in out.items() if k != "done"}
    # correctness of the host round trip: same bits as the device-resident pass
    ref_bits = w.step(0)["bits"].double().cpu()
    torch.cuda.synchronize()
    # (chunked launches group the fp32 partials differently; NOISE launches draw fresh noise per run)
    comparable = not c.training
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
    # the same bytes with no kernels at all, both directions at once on the pipeline's own streams and buffers, all ranks
    # together: the ceiling the host links / host memory set for this step (what e2e can reach at most)
    slot = hp.slots[0]
    if world > 1:
        dist.barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    hp.s_h2d.wait_stream(cur)
    hp.s_d2h.wait_stream(cur)
    c0.record()
    hp.s_h2d.wait_event(c0)
    hp.s_d2h.wait_event(c0)
    dev_out = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in slot["h_out"].items()}
    for _ in range(steps):
        with torch.cuda.stream(hp.s_h2d):
            for k in ("y", "mu", "sigma", "z"):
                slot["d_in"][k].copy_(host[k], non_blocking=True)
        with torch.cuda.stream(hp.s_d2h):
            for k, v in slot["h_out"].items():
                v.copy_(dev_out[k], non_blocking=True)
    cur.wait_stream(hp.s_h2d)
    cur.wait_stream(hp.s_d2h)
    c1.record()
    torch.cuda.synchronize()
    cms = c0.elapsed_time(c1)
    if world > 1:
        t = torch.tensor([cms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cms = float(t.item())
    ceiling = {"ms_per_step": cms / steps, "value": global_elems / (cms / steps * 1e-3) / 1e6,
               "h2d_gb_per_s_per_rank": hp.h2d_bytes / (cms / steps * 1e-3) / 1e9,
               "what": "the step's H2D and D2H copies alone (no kernels), all ranks at once, max over ranks"}
    return {"value": global_elems / (ms / steps * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": hp.h2d_bytes,
            "d2h_bytes_per_step": hp.d2h_bytes, "ms_per_step": ms / steps, "steps": steps, "bytes_are": "per rank",
            "copy_ceiling": ceiling,
            "what": f"pinned host y/mu/sigma/z -> H2D -> 1+5{'+1' if packed_slots else ''} launches -> D2H {'/'.join(hp.out_names)}, "
                    f"{len(hp.ranges)} image chunks pipelined on 3 streams, batches double-buffered"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
