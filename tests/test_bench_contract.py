"""bench.py prints one JSON line with the contract's keys (reference arm on CPU here; our arm on the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "cpu_baseline", "e2e"}


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-seconds", "1"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 100 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_c_port"]["value"] > d["value"]  # the C port is reported next to the Python loop


@pytest.mark.gpu
def test_our_arm_line():
    d = _run(["--steps", "4", "--warmup", "3", "--envs-per-gpu", "262144", "--e2e-steps", "2", "--cpu-seconds", "1", "--sustained-seconds", "0.2"])
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "sustained"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["gpu_launches"] == 4 and d["scaling"] == "weak" and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.05 < r["frac"] < 1.2
    assert abs(d["value"] - 262144 * 4 / (d["ms_per_step"] * 4e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * 262144 and e["d2h_bytes_per_step"] == 262144 * (576 + 10) and 0 < e["value"] < d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
