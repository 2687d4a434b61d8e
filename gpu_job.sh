mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_n1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_n1.log)
tail -n 3 gpurun_out/pytest_gpu_n1.log
for s in 128 256; do
python bench.py --size $s --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/diet_s$s.json 2> gpurun_out/diet_s$s.err
python - <<PY
import json
d=json.loads(open("gpurun_out/diet_s$s.json").read().strip().splitlines()[-1])
print("diet s$s", round(d["value"]/1e9,3), "G  ms", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["roofline"]["per_kernel_ms"].items()})
PY
done
