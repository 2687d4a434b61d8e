mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log)
python bench.py --no-cpu-baseline > gpurun_out/bench_v4_default.json 2> gpurun_out/bench_v4_default.err
for v in k1_2_k3_3; do
  LULESH_B200_LIB=$PWD/build/variants/lib_$v.so python bench.py --no-cpu-baseline > gpurun_out/bench_v4_$v.json 2> gpurun_out/bench_v4_$v.err
done
LULESH_B200_LIB=$PWD/build/variants/lib_k1_2_k3_3.so python bench.py --size 256 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v4_s256.json 2>&1
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_v4_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"]/1e9,3), {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms"].items()})
    except Exception as e: print(f, "ERR", e, open(f).read()[-300:])
PY
