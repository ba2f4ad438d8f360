"""bench.py prints ONE JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_runs_on_the_host_cores():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "sv_particle_steps_per_sec" and d["unit"] == "particle-steps/s"
    assert d["value"] > 1e5 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "4096 proposals x 1024 particles" in d["config"]["workload"]


@pytest.mark.gpu
def test_gpu_arm_contract():
    d = _run(["--steps", "1", "--warmup", "1", "--proposals", "296", "--T", "256", "--no-pmmh"])
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "layout"} <= set(d)
    assert d["value"] > 1e9 and d["e2e"]["value"] > 1e8 and d["gpu_launches"] >= 2
    assert d["e2e"]["h2d_bytes_per_step"] == 296 * 3 * 8 and d["e2e"]["d2h_bytes_per_step"] == 296 * 8
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert 0 < d["roofline"]["frac"] < 1.2
    assert d["cpu_baseline"]["value"] > 1e5 and d["cpu_baseline"]["cores"] >= 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # the bench checks itself: a sample of the timed filters against the oracle, and no leg may have failed
    assert d["parity_sample"]["checked"] == 4 and d["parity_sample"]["bit_identical"] == 4
    assert d["legs_ok"] is True
    assert list(d)[-2:] == ["pmmh", "spilled_filter"]
