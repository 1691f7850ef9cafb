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
    assert d["config"]["workload"].startswith("tcm64_kodak768x512_b24")


def test_reference_arm_other_ranks_exit_quietly():
    assert _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []


@pytest.mark.gpu
def test_default_arm_line_on_gpu():
    lines = _run(["--steps", "12", "--warmup", "3", "--cpu-images", "1", "--no-whole-y", "--no-training-kernels"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["n_gpus"] == 1 and d["gpu_launches"] == 6 * 12
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0.2 < r["frac"] < 1.0 and r["kernel"] == "gc_fwd_kernel"
    assert abs(r["achieved"] / r["peak"] - r["frac"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 143327232 and d["e2e"]["d2h_bytes_per_step"] == 94372032
    assert d["e2e_slots"]["d2h_bytes_per_step"] < 0.6 * d["e2e"]["d2h_bytes_per_step"]
    assert d["e2e"]["value"] < d["value"] and d["clocks"]["sm_max_mhz"]
