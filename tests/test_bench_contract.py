"""The bench contract, checked without a GPU: argument defaults, the reference arm's line (it runs on the host cores), and
the keys of the committed lines of both arms (profiles/r2/bench_final.json is what `python bench.py` printed on a B200)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches")


def _line(path):
    with open(path) as f:
        return json.loads([ln for ln in f if ln.startswith("{")][0])


def test_defaults_follow_the_contract(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert a.gpus == 1 and a.warmup >= 3 and a.steps >= 1 and a.impl == "b200"


def test_committed_lines_carry_every_contract_key():
    d = _line(os.path.join(ROOT, "profiles", "r2", "bench_final.json"))
    for k in REQUIRED + ("roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == "vehicle-steps/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= 1.02 * d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ref = _line(os.path.join(ROOT, "profiles", "r2", "bench_ref_final.json"))
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["gpu_launches"] == 0 and ref["cpu_baseline"]["value"] == ref["value"]


def test_reference_arm_runs_on_the_host_and_other_ranks_stay_silent():
    """`bench.py --impl reference`: the reference's own CPU implementation (oracle/_ref when built, else the port) on the host
    cores; under torchrun only rank 0 works and prints."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-seconds", "2"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 1e5 and d["cpu_baseline"]["kind"] in ("reference", "port")
