mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --size 256 --no-cpu-baseline > gpurun_out/bench_n2_s256.json 2> gpurun_out/bench_n2_s256.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_n2*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]/1e9,3), d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms"].items()}, d["e2e"]["value"]/1e9, d["gpu_launches"])
    except Exception as e: print(f, "ERR", e, open(f).read()[-300:])
PY
