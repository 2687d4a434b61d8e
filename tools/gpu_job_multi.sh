#!/bin/bash
# Multi-GPU check of a round-2 build (gpurun --gpus N): multi-rank parity tests, then the bench the way the
# driver launches it (torchrun, one rank per GPU) at the default size and at -s 128 per GPU.
N=${1:-2}; out=gpurun_out/${2:-r02_multi$N}; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
(timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x --timeout 150 --timeout-method=thread > $out/pytest_multi.log 2>&1; echo "pytest exit $?" >> $out/pytest_multi.log)
tail -n 4 $out/pytest_multi.log
run() { timeout -k 10 ${STEP_TIMEOUT:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 20 --warmup 5 > $out/bench_n$N.json 2> $out/bench_n$N.err; echo "bench rc $?"
if [ "$N" = 8 ]; then   # the 4-rank layout on the same box: parity + weak numbers, no extras
   N=4; run --steps 20 --warmup 5 --no-extras > $out/bench_n4.json 2> $out/bench_n4.err; echo "bench n4 rc $?"; N=8
   python - $out <<'PY'
import json, sys
try:
    d = json.load(open(f"{sys.argv[1]}/bench_n4.json"))
    p = d.get("parity") or {}
    print("N=4 s256: %.3f Gz/s %.3f ms/step parity %s" % (d["value"]/1e9, d["ms_per_step"], p.get("status")),
          {m: r.get("ok") for m, r in (p.get("modes") or {}).items()})
except Exception as e:
    print("bench_n4 unreadable", e)
PY
fi
run --size 128 --steps 200 --warmup 10 --no-extras --no-parity > $out/bench_n${N}_s128.json 2> $out/bench_n${N}_s128.err; echo "bench s128 rc $?"
LULESH_B200_HALO=nccl run --size 128 --steps 200 --warmup 10 --no-extras --no-parity > $out/bench_n${N}_s128_nccl.json 2> $out/bench_n${N}_s128_nccl.err; echo "bench s128 nccl rc $?"
timeout 300 python bench.py --size 128 --steps 200 --warmup 10 --no-extras --no-cpu-baseline > $out/bench_n1_s128.json 2> $out/bench_n1_s128.err
python - $out $N <<'PY'
import json, sys
out, N = sys.argv[1], sys.argv[2]
def load(f):
    try: return json.load(open(f"{out}/{f}"))
    except Exception as e: print(f, "unreadable", e); return None
d = load(f"bench_n{N}.json")
if d:
    print("N=%s s256: %.3f Gz/s  %.3f ms/step  halo=%s" % (N, d["value"]/1e9, d["ms_per_step"], d["halo"]))
    print(" timeline", {k: round(v, 4) for k, v in (d.get("timeline_ms") or {}).items()})
    print(" nccl", d.get("nccl_fallback"))
    p = d.get("parity") or {}
    print(" parity", p.get("status"), {m: {k: v for k, v in r.items() if k in ("ok","cycles","e0_rel_err","shared_nodes_bit_identical","scalars_bit_identical_on_all_ranks","halo")} for m, r in (p.get("modes") or {}).items()})
    for k, v in (d.get("extras") or {}).items():
        print(" ", k, round(v["value"]/1e9, 3), "Gz/s", round(v["ms_per_step"], 4), "ms")
a, b, c = load("bench_n1_s128.json"), load(f"bench_n{N}_s128.json"), load(f"bench_n{N}_s128_nccl.json")
if a and b:
    print("s128 weak: N=1 %.4f ms, N=%s p2p %.4f ms (eff %.3f)" % (a["ms_per_step"], N, b["ms_per_step"], a["ms_per_step"]/b["ms_per_step"]))
    print(" timeline", {k: round(v, 4) for k, v in (b.get("timeline_ms") or {}).items()})
if a and c:
    print("           N=%s nccl %.4f ms (eff %.3f) halo=%s" % (N, c["ms_per_step"], a["ms_per_step"]/c["ms_per_step"], c["halo"]))
    print(" timeline", {k: round(v, 4) for k, v in (c.get("timeline_ms") or {}).items()})
PY
