mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log)
tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --no-cpu-baseline > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err
python bench.py --size 256 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v6_s256.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_v6*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"]/1e9,3), {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e9,3))
    except Exception as e: print(f, "ERR", e, open(f).read()[-300:])
PY
