"""bench.py's output contract: exactly ONE JSON line on stdout with the keys the driver reads.  The reference arm
runs on the CPU (here, on a small grid); the CUDA arm needs a device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(args, timeout=900):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         cwd=ROOT, timeout=timeout)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]              # one line, nothing else on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    j = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--grid", "12,10,8", "--nodes", "4"])
    assert j["impl"] == "reference" and BASE_KEYS <= set(j)
    assert j["unit"] == "edge-updates/s" and j["higher_is_better"] is True and j["dtype"] == "f64"
    assert j["steps"] >= 1 and j["value"] > 0 and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["gpu_launches"] == 0


@pytest.mark.gpu
def test_cuda_arm_line():
    j = _run(["--steps", "4", "--warmup", "3", "--grid", "24,20,16", "--nodes", "8", "--e2e-steps", "3"])
    assert "impl" not in j or j["impl"] != "reference"
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches"} <= set(j)
    assert j["n_gpus"] == 1 and j["steps"] == 4 and j["warmup"] == 3 and j["value"] > 0 and j["gpu_launches"] > 0
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "traffic" in r
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < j["value"] * 1.05
    c = j["cpu_baseline"]
    assert c["value"] > 0 and c["cores"] >= 1 and c["kind"] in ("port", "reference") and c["sample"]
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert len(j["config"]["iter_ms"]) == 7 and j["config"]["objective_trace"]["digest"]
