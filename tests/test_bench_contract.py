"""bench.py contract on the CPU: the reference arm prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-cells", "2e5"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "FDTD Mcell-updates/s" and d["unit"] == "Mcell/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["value"] > 0
    assert d["config"]["workload"] == "patch100m"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "cells" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_own_arm_prints_one_json_line_with_every_record():
    """the GPU arm on a reduced mesh: one JSON line with the keys the driver and the judge read (value, roofline with the
    real-traffic fraction, e2e through the public API, the seeded per-cell-coefficient record, gpu_launches, clocks)"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "6", "--warmup", "3", "--cells", "3e6",
                          "--no-strong", "--no-latency", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["metric"] == "FDTD Mcell-updates/s" and d["unit"] == "Mcell/s" and d["n_gpus"] == 1 and d["steps"] == 6
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0 and d["higher_is_better"] is True
    assert d["config"]["workload"] == "patch100m" and d["config"]["timed_region"]["sampling_launch_sets"] >= 1
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["achieved"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] > 0 and 0 < r["actual_frac"] < 1.2
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["operator_restored_mismatches"] == [0, 0, 0, 0]
    assert d["seeded"]["value"] > 0 and d["seeded"]["roofline"]["kernel_ms"]
    assert "sm_mhz" in d["clocks"]
