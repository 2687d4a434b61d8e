mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/pytest_multi_p2p.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi_p2p.log)
tail -4 gpurun_out/pytest_multi_p2p.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 200 --no-cpu-baseline > gpurun_out/n2_p2p.json 2> gpurun_out/n2_p2p.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n2_p2p.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["config"].get("halo"), d["n_gpus"], round(d["value"]/1e9,3), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e9,3), d["gpu_launches"])
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-800:])
PY
