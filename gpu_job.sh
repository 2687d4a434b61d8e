mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_force|k_kinematics|k_material|k_node" -s 40 -c 4 -o gpurun_out/prof_s128_v8 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v8.log 2>&1
python bench.py --steps 200 --no-cpu-baseline > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err
python bench.py --size 256 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v8_s256.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_v8*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"]/1e9,3), {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e9,3))
    except Exception as e: print(f, "ERR", e, open(f).read()[-300:])
PY
