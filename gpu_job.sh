mkdir -p gpurun_out
Q='kernels_one_by_one or cycles_against_oracle or step_equals or deterministic or error_codes'
run_quick() { # name, env...
  name=$1; shift
  (env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --timeout 60 -k "$Q" > gpurun_out/quick_$name.log 2>&1; echo "pytest exit $?" >> gpurun_out/quick_$name.log)
  echo "quick $name: $(tail -n 2 gpurun_out/quick_$name.log | tr '\n' ' ')"
}
run_quick nofuse LULESH_B200_FUSE=0
run_quick fused X=1
run_quick lag4 LULESH_B200_LIB=$PWD/lulesh_b200/lib/liblulesh_b200_lag4.so
run_quick fc0 LULESH_B200_LIB=$PWD/lulesh_b200/lib/liblulesh_b200_fc0.so
if grep -L "pytest exit 0" gpurun_out/quick_*.log | grep -q .; then echo "SOME QUICK TESTS FAILED"; grep -L "pytest exit 0" gpurun_out/quick_*.log; fi
bench() { # name size env...
  name=$1; s=$2; shift 2
  env "$@" timeout 300 python bench.py --size $s --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/b_${name}_s$s.json 2> gpurun_out/b_${name}_s$s.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_${name}_s$s.json").read().strip().splitlines()[-1])
    print("$name s$s", round(d["value"]/1e9,3), "G  ms", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["roofline"]["per_kernel_ms"].items()})
except Exception as e:
    print("$name s$s FAILED", e)
PY
}
bench nofuse 128 LULESH_B200_FUSE=0
bench fused 128 X=1
bench lag4 128 LULESH_B200_LIB=$PWD/lulesh_b200/lib/liblulesh_b200_lag4.so
bench fc0 128 LULESH_B200_LIB=$PWD/lulesh_b200/lib/liblulesh_b200_fc0.so
bench nofuse 256 LULESH_B200_FUSE=0
bench lag4 256 LULESH_B200_LIB=$PWD/lulesh_b200/lib/liblulesh_b200_lag4.so
