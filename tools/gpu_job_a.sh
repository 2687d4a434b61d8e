#!/bin/bash
# 1-GPU check of a round-2 build: full single-GPU parity suite, then both bench arms the way the driver calls them.
out=gpurun_out/${1:-r02_a}; mkdir -p $out
(timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log)
tail -n 5 $out/pytest_gpu.log
(python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log); tail -n 2 $out/smoke.log
python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc $?"
python bench.py --impl reference --steps 20 --warmup 5 > $out/ref_n1.json 2> $out/ref_n1.err; echo "ref rc $?"
python - $out <<'PY'
import json, sys
out = sys.argv[1]
d = json.load(open(f"{out}/bench_n1.json"))
print("value %.3f Gz/s  ms/step %.3f  e2e %.3f Gz/s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9))
print("per kernel", {k: round(v, 4) for k, v in d["roofline"]["per_kernel_ms"].items()}, "step frac", d["roofline"]["step"])
for k, v in (d.get("extras") or {}).items():
    print(k, {kk: (round(vv, 5) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "workload"})
print("cpu", d["cpu_baseline"])
r = json.load(open(f"{out}/ref_n1.json"))
print("ref", r["value"], r["steps"], r["warmup"], r["cpu_baseline"]["sample"], "same_config", r["config"] == d["config"])
PY
