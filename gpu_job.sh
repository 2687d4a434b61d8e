mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log)
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --no-cpu-baseline > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_v11*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"]/1e9,3), {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms"].items()})
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-300:])
PY
