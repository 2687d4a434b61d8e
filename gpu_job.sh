mkdir -p gpurun_out
python bench.py > gpurun_out/final_s128.json 2> gpurun_out/final_s128.err
python bench.py --size 256 --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/final_s256.json 2> gpurun_out/final_s256.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 50 --csv --log-file gpurun_out/final_launches_s128.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_force|k_kinematics|k_material|k_node" -s 40 -c 4 -f -o gpurun_out/prof_s128_final python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/final_s128.json","gpurun_out/final_s256.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, round(d["value"]/1e9,4), round(d["ms_per_step"],4), {k:round(v,4) for k,v in r["per_kernel_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e9,3), "roof", round(r["frac"],3), "step", round(r["step"]["frac"],3), d.get("cpu_baseline"), d.get("clocks"))
PY
