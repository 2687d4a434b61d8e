#!/bin/bash
# Round-2 kernel sweep on one B200 (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash tools/sweep_job.sh <outdir> <tag> [<tag> ...]'
# For every variant library lulesh_b200/lib/variants/lib_<tag>.so ("default" = the product
# library): the per-kernel fixture test + oracle cycles, then bench lines at -s 128 (and
# -s 256 when SWEEP_S256=1).
out=gpurun_out/$1; shift
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt
for tag in "$@"; do
   if [ "$tag" = default ]; then unset LULESH_B200_LIB; else export LULESH_B200_LIB=$PWD/lulesh_b200/lib/variants/lib_$tag.so; fi
   (timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "one_by_one or against_oracle or goldens" > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log)
   tail -n 2 $out/pytest_$tag.log | tr '\n' ' '; echo " <= $tag"
   LULESH_B200_VERBOSE=1 timeout 300 python bench.py --size 128 --steps 200 --warmup 10 --no-cpu-baseline > $out/b_${tag}_s128.json 2> $out/b_${tag}_s128.err
   if [ "$SWEEP_S256" = 1 ]; then
      timeout 300 python bench.py --size 256 --steps 40 --warmup 5 --no-cpu-baseline > $out/b_${tag}_s256.json 2> $out/b_${tag}_s256.err
   fi
   python - "$out" "$tag" <<'PY'
import json, sys
out, tag = sys.argv[1:3]
for size in (128, 256):
    try:
        d = json.load(open(f"{out}/b_{tag}_s{size}.json"))
    except Exception:
        continue
    pk = d["roofline"]["per_kernel_ms"]
    print(f"{tag:10s} s{size}: {d['value']/1e9:.3f} Gz/s  {d['ms_per_step']*1e3:.1f} us/cycle  " +
          "  ".join(f"{k[:6]}={v*1e3:.1f}" for k, v in pk.items()))
PY
done
