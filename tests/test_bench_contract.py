"""CPU: the measurement contract of bench.py that can be checked without a GPU — the reference arm prints
ONE JSON line with the agreed keys, ranks other than 0 stay silent, and the clocks parser reads nvidia-smi's
CSV as documented."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=e, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "patches/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["metric"].startswith("1024x1024 patches/sec") and d["config"]["workload"] == "cell_1024x1024_b8_k11_5step"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["scaling"] == "weak"
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e2e == {"value": d["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["ms_per_step"] * d["value"] / 1e3 - 1.0) < 1e-6          # one patch per step


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_clocks_parser():
    sys.path.insert(0, ROOT)
    import bench
    c = bench.Clocks(0)
    c.rows = [(10.0, "1965, 1965, 640.1, Not Active, Not Active, Not Active, Not Active"),
              (10.02, "1950, 1965, 700.0, Not Active, Not Active, Not Active, Active"),
              (10.04, "[N/A], 1965, 1, x, x, x, x"),
              (99.0, "300, 1965, 80.0, Active, Active, Active, Active")]       # outside the window
    s = c.summary(9.9, 10.1)
    assert s == {"sm_mhz": 1957.5, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 2}
    assert bench.Clocks(0).summary(0, 1) == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
