mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q --timeout 120 > gpurun_out/pytest_multi_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi_final.log)
tail -n 3 gpurun_out/pytest_multi_final.log
python bench.py --no-cpu-baseline --steps 200 > gpurun_out/scale2_n1.json 2> gpurun_out/scale2_n1.err
for n in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 200 --no-cpu-baseline > gpurun_out/scale2_n$n.json 2> gpurun_out/scale2_n$n.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29750 bench.py --gpus 8 --size 256 --steps 50 --no-cpu-baseline > gpurun_out/scale2_n8_s256.json 2> gpurun_out/scale2_n8_s256.err
LULESH_B200_HALO=nccl timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29760 bench.py --gpus 8 --steps 200 --no-cpu-baseline > gpurun_out/scale2_n8_nccl.json 2> gpurun_out/scale2_n8_nccl.err
(timeout 120 ./lulesh_b200/bin/lulesh_b200 --gpus 8 --global 384 -i 100 > gpurun_out/driver_cfg4_8gpu.txt 2>&1; echo "exit $?" >> gpurun_out/driver_cfg4_8gpu.txt)
(timeout 120 ./lulesh_b200/bin/lulesh_b200 --gpus 1 -s 384 -i 100 > gpurun_out/driver_cfg4_1gpu.txt 2>&1; echo "exit $?" >> gpurun_out/driver_cfg4_1gpu.txt)
(timeout 300 ./lulesh_b200/bin/lulesh_b200 --gpus 8 -s 320 -i 30 > gpurun_out/driver_cfg5_8gpu.txt 2>&1; echo "exit $?" >> gpurun_out/driver_cfg5_8gpu.txt)
grep -E "Iteration|Origin|FOM|Elapsed|exit|MPI" gpurun_out/driver_cfg4_8gpu.txt gpurun_out/driver_cfg4_1gpu.txt gpurun_out/driver_cfg5_8gpu.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale2_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["config"].get("halo"), d["n_gpus"], round(d["value"]/1e9,3), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["per_kernel_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e9,3))
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-400:])
PY
