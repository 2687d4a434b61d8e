"""bench.py contract pieces that do not need a GPU: the reference arm (`--impl reference`)
prints one JSON line with the agreed keys; helpers behave."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "12",
                        "--steps", "20", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "zones/s" and d["higher_is_better"] is True
    assert d["metric"] == "LULESH FOM (zone-cycles/s)" and d["dtype"] == "f64" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "zones/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("-s 12 ")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_helpers():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.reference_cycle_budget(128, 10**9, 20.0) >= 2
    assert bench.reference_cycle_budget(30, 50, 60.0) == 50
    peak, src = bench.measured_peak_gbs()
    assert 3000 < peak < 9000 and ("measured" in src or "fallback" in src)
    assert sum(v for k, v in bench.B_ALG.items()) == bench.B_ALG_STEP
