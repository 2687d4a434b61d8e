mkdir -p gpurun_out
LULESH_B200_LIB=$PWD/build/variants/lib_k1park2.so python bench.py --steps 200 --no-cpu-baseline > gpurun_out/bench_v13_park2.json 2> gpurun_out/bench_v13_park2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_v13*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"]/1e9,3), {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms"].items()})
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-300:])
PY
