mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_n1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_n1.log)
tail -n 3 gpurun_out/pytest_gpu_n1.log
bench() { # name size
  name=$1; s=$2
  lib=$PWD/lulesh_b200/lib/liblulesh_b200$name.so
  LULESH_B200_LIB=$lib timeout 300 python bench.py --size $s --steps 150 --warmup 10 --no-cpu-baseline > gpurun_out/sw${name}_s$s.json 2> gpurun_out/sw${name}_s$s.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sw${name}_s$s.json").read().strip().splitlines()[-1])
    print("variant[$name] s$s", round(d["value"]/1e9,3), "G  ms", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["roofline"]["per_kernel_ms"].items()})
except Exception as e:
    print("variant[$name] s$s FAILED", e)
PY
}
for v in "" _k2_128 _k2_full _mat256 _mat64 _mat8 _k3_4; do bench "$v" 128; done
bench "" 256; bench _mat8 256; bench _k2_full 256
