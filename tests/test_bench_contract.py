"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the contract's keys; under a
multi-rank launch only rank 0 prints.  GPU: the default arm's line carries roofline / cpu_baseline / e2e / clocks."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-images", "1"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "entropy_model_latent_melem_per_s" and d["unit"] == "Melem/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "tcm128_kodak768x512_b64_likelihood_bpp" and d["scaling"] == "strong"
    # the reference arm steps over a bounded sample but names the SAME workload object as our arm
    import bench

    assert d["config"] == bench.workload_config(bench.synthetic.CONFIGS[3], 1, "strong")


def test_reference_arm_other_ranks_exit_quietly():
    assert _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []


@pytest.mark.gpu
def test_default_arm_line_on_gpu():
    lines = _run(["--steps", "24", "--warmup", "3", "--cpu-images", "1", "--no-training-kernels", "--legs", "2"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["n_gpus"] == 1 and d["gpu_launches"] == 6 * 24 and d["scaling"] == "strong"
    assert d["config"]["workload"] == "tcm128_kodak768x512_b64_likelihood_bpp" and d["config"]["images_per_gpu"] == 64
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0.2 < r["frac"] < 1.0 and r["kernel"] == "gc_fwd_kernel"
    assert abs(r["achieved"] / r["peak"] - r["frac"]) < 1e-9 and r["bytes_per_elem"] == 20 and r["elems_per_launch"] == 64 * 98304
    # the kernel alone (one dependent chain) is reported beside the pattern of the timed region (3 batches in flight)
    assert 0.2 < r["single_chain"]["frac"] <= r["frac"] * 1.05 and "in flight" in r["launch"]
    assert 5 * r["us_per_launch"] * 1e-3 <= d["ms_per_step"] * 1.02          # the GC launches fit inside the step
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    # config 3 reads back the per-image bits only; its inputs are 64 images of y / mu / sigma / z
    assert d["e2e"]["h2d_bytes_per_step"] == 64 * (3 * 491520 + 18432) * 4 and d["e2e"]["d2h_bytes_per_step"] == 64 * 8
    assert d["e2e"]["value"] < d["value"] and d["clocks"]["sm_max_mhz"]
    # the other BASELINE configs carry their own roofline objects; config 2 is the compress-path one (packed slots home)
    leg = d["per_config"]["2"]
    assert leg["roofline"]["bytes_per_elem"] == 28 and leg["roofline"]["elems_per_launch"] == 24 * 98304
    assert 0.2 < leg["whole_y"]["roofline"]["frac"] < 1.0
    assert leg["e2e"]["h2d_bytes_per_step"] == 143327232 and leg["e2e"]["d2h_bytes_per_step"] < 0.6 * 94372032
    assert d["per_config"]["3"]["roofline"] == d["roofline"] and d["legs"]["2"]["scaling"] == "strong"


@pytest.mark.gpu
def test_simulated_shard_line_on_gpu():
    """--shard-of 8: rank 0's share of the 8-GPU strong-scaling job on one GPU (development aid)."""
    lines = _run(["--steps", "24", "--warmup", "3", "--shard-of", "8", "--no-e2e", "--no-whole-y", "--no-cpu-baseline", "--legs", "none"])
    d = json.loads(lines[0])
    assert d["config"]["images_per_gpu"] == 8 and d["config"]["simulated_shard_of"] == 8 and d["value"] > 0


def test_graph_and_branch_shaping_is_even_for_any_step_count():
    """bench.balanced_group / pick_chains: the timed region is cut into equal graphs, and the batches in flight deal a
    graph's steps evenly onto the branches (the driver runs --steps 20; the default is 200)."""
    import bench

    c3, c5 = bench.synthetic.CONFIGS[3], bench.synthetic.CONFIGS[5]
    assert bench.balanced_group(200, 48) == 40 and bench.balanced_group(20, 48) == 20 and bench.balanced_group(192, 48) == 48
    assert bench.balanced_group(1, 48) == 1 and bench.balanced_group(49, 48) == 25
    for steps in (1, 7, 20, 24, 100, 192, 200):
        g = bench.balanced_group(steps, 48)
        assert 1 <= g <= 48 and -(-steps // g) == -(-steps // 48)          # as few graphs as the cap allows, none nearly empty
    assert bench.pick_chains(c3, 64, 40) == 4 and bench.pick_chains(c3, 64, 20) == 4          # full batch: 4 in flight
    assert bench.pick_chains(c3, 8, 20) == 5 and bench.pick_chains(c3, 8, 40) == 5 and bench.pick_chains(c3, 8, 48) == 6
    assert bench.pick_chains(c5, 32, 20) == 5 and bench.pick_chains(c5, 256, 20) == 4
