mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run() { # name n env...
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then
    env "$@" timeout 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale2_$name.json 2> gpurun_out/scale2_$name.err
  else
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $n --no-cpu-baseline > gpurun_out/scale2_$name.json 2> gpurun_out/scale2_$name.err
  fi
  python - <<PY
import json
try:
    txt=open("gpurun_out/scale2_$name.json").read().strip().splitlines()
    d=json.loads(txt[-1])
    print("$name", "lines", len(txt), round(d["value"]/1e9,3), "G  ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]/1e9,3), d["config"].get("halo"), d["clocks"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/scale2_$name.err").read()[-500:])
PY
}
run n2_a 2 X=1
run n8 8 X=1
run n2_b 2 X=1
run n4 4 X=1
run n2_nccl 2 LULESH_B200_HALO=nccl
run n1 1 X=1
