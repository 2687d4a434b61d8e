mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log)
tail -n 3 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --no-cpu-baseline > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err
python bench.py --size 256 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v12_s256.json 2>&1
python bench.py --size 256 --regions 16 --balance 1 --cost 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v12_cfg3.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_v12*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"]/1e9,3), {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms"].items()})
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-300:])
PY
