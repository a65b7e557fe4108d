"""The bench.py output contract (one JSON line on stdout with the keys the driver reads).  CPU tier: the reference arm on a
tiny time budget; GPU tier: the B200 arm with a handful of steps."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"bench.py must print exactly one line on stdout, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 300)
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "afterstates/sec" and d["unit"] == "afterstates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["config"]["envs_total"] == 1 << 20 and "workload" in d["config"]
    staged = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "game", "tetris.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "reference_sample" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_b200_arm_line(gpu):
    d = _run(["--steps", "4", "--warmup", "3", "--no-dqn"], 900)
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "e2e_40slot", "distinct_form", "collective_us", "pcie"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["gpu_launches"] == 4 and d["scaling"] == "weak" and d["vs_baseline"] is None
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["traffic"] is None or rf["traffic"] > 0          # null unless profiles/r02_ncu_current.json matches this build
    n = 1 << 20
    assert d["e2e_40slot"]["h2d_bytes_per_step"] == 2 * n and d["e2e_40slot"]["d2h_bytes_per_step"] == 163 * n
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * n and 90 * n < d["e2e"]["d2h_bytes_per_step"] < 110 * n
    assert 0 < d["e2e_40slot"]["value"] < d["e2e"]["value"] < d["value"]
    assert 22 < d["distinct_form"]["words_per_env"] < 24.5
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["config"] == dict(d["config"]) and "model" not in d["config"]
