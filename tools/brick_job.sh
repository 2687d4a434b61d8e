mkdir -p gpurun_out
Q='kernels_one_by_one or cycles_against_oracle or step_equals or deterministic or device_side_setup or error_codes'
(timeout 120 python -m pytest tests/test_gpu_parity.py -x -q --timeout 60 -k "$Q" > gpurun_out/quick_default.log 2>&1; echo "pytest exit $?" >> gpurun_out/quick_default.log)
echo "default: $(tail -n 2 gpurun_out/quick_default.log | tr '\n' ' ')"
(LULESH_B200_BRICK=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q --timeout 60 -k "$Q" > gpurun_out/quick_brick.log 2>&1; echo "pytest exit $?" >> gpurun_out/quick_brick.log)
echo "brick: $(tail -n 2 gpurun_out/quick_brick.log | tr '\n' ' ')"
grep -E "^E |Error|FAILED" gpurun_out/quick_brick.log | head -8
for mode in 1 0; do
LULESH_B200_BRICK=$mode timeout 100 python bench.py --size 128 --steps 150 --warmup 10 --no-cpu-baseline > gpurun_out/brick${mode}_s128.json 2> gpurun_out/brick${mode}_s128.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/brick${mode}_s128.json").read().strip().splitlines()[-1])
    print("brick=$mode s128", round(d["value"]/1e9,3), "G  ms", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["roofline"]["per_kernel_ms"].items()})
except Exception as e:
    print("brick=$mode FAILED", e); print(open("gpurun_out/brick${mode}_s128.err").read()[-300:])
PY
done
