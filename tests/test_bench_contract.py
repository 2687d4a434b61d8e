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


def test_reference_arm_never_loads_the_product_library(tmp_path):
    """The arm must run with no product library at all, and echo steps/warmup/config like the b200 arm."""
    env = dict(os.environ, LULESH_B200_LIB=str(tmp_path / "missing.so"))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "10",
                        "--steps", "7", "--warmup", "4"], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][0])
    assert d["steps"] == 7 and d["warmup"] == 4
    sys.path.insert(0, ROOT)
    import bench
    args = bench.argparse.Namespace(size=10, glob=0, regions=11, balance=1, cost=1, steps=7, warmup=4)
    assert d["config"] == bench.workload_config(args, 1)     # same dict as the b200 arm prints
    assert "lulesh_b200" not in " ".join(d["cpu_baseline"]["sample"].split("lulesh_omp"))


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
    assert bench.decompose(1) == (1, 1, 1) and bench.decompose(2) == (1, 1, 2)
    assert bench.decompose(4) == (1, 2, 2) and bench.decompose(8) == (2, 2, 2) and bench.decompose(27) == (3, 3, 3)
    import lulesh_b200 as lb
    for n in (1, 2, 4, 8, 27):
        assert bench.decompose(n) == lb.decompose(n)          # the restated rule is the library's
    ns = bench.argparse.Namespace(size=0, glob=384, regions=11, balance=1, cost=1, steps=5, warmup=3)
    assert bench.rank_sizes(ns, 4) == ((1, 2, 2), (384, 192, 192))
    assert bench.reference_sample_size(256, 8, 10 ** 6, 1.0) <= 32
