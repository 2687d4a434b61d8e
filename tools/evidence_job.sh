#!/bin/bash
# One-GPU evidence job behind the files under profiles/ (run on a B200 box from the repo root):
#   gpurun --timeout 1700 -- 'bash tools/evidence_job.sh r02'
# then, back in the container:
#   python profiles/summarize_ncu.py gpurun_out/r02/prof_s256.ncu-rep r02_s256 --size 256
#   python profiles/summarize_ncu.py gpurun_out/r02/prof_s128.ncu-rep r02_s128 --size 128
# Every step runs under its own timeout: a hung step costs minutes, not the box.
tag=${1:-r02}; out=gpurun_out/$tag; mkdir -p $out
free -g | head -n 2 > $out/host.txt; nproc >> $out/host.txt; lscpu | grep "Model name" >> $out/host.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv >> $out/host.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log)
(timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log)
tail -n 2 $out/smoke.log; tail -n 3 $out/pytest_gpu.log
# the two arms exactly as the driver calls them, then the default line with more steps
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_driver_like.json 2> $out/bench_driver_like.err
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
# launch list (kernel shares of the step) and one full capture per kernel; numbers printed under ncu are never bench values
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 50 --csv --log-file $out/launches_s256.csv \
   python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force|k_kinematics|k_material|k_node" -s 16 -c 4 -f -o $out/prof_s256 \
   python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $out/ncu_s256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force|k_kinematics|k_material|k_node" -s 40 -c 4 -f -o $out/prof_s128 \
   python bench.py --size 128 --steps 12 --warmup 3 --no-extras --no-cpu-baseline > $out/ncu_s128.log 2>&1
# R3 evidence: FP64-pipe instructions of the material kernel against the EOS repetition count
# (-r 1 -c C gives every element rep = 1 + C, lulesh.cc:2393-2400; -c 0 -> 1, 1 -> 2, 8 -> 9, 19 -> 20)
B=./lulesh_b200/bin/lulesh_b200
for c in 0 1 8 19; do
   timeout 300 ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none \
      -k regex:k_material -s 2 -c 1 --csv --log-file $out/r3_rep$((c + 1)).csv $B -s 64 -i 4 -r 1 -c $c -q > /dev/null 2>&1
done
# sanitizers on a small problem, host and device-side setup
for tool in memcheck racecheck initcheck; do
  (echo "# compute-sanitizer --tool $tool $B -s 12 -i 12 -r 5 -c 2 [--device-setup] -q";
   timeout 280 compute-sanitizer --tool $tool $B -s 12 -i 12 -r 5 -c 2 -q 2>&1 | tail -n 4;
   timeout 280 compute-sanitizer --tool $tool $B -s 12 -i 12 -r 5 -c 2 --device-setup -q 2>&1 | tail -n 4;
   echo "$tool exit $?") > $out/sanitizer_$tool.log 2>&1
done
ls -la $out | tail -n 30
