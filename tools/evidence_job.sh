#!/bin/bash
# One-GPU evidence job behind the files under profiles/ (run on a B200 box from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/evidence_job.sh'
# then, back in the container:
#   python profiles/summarize_ncu.py gpurun_out/prof_s128_final.ncu-rep r01_s128_final --size 128
#   python profiles/summarize_ncu.py gpurun_out/prof_s256_final.ncu-rep r01_s256_final --size 256
mkdir -p gpurun_out
free -g | head -n 2 > gpurun_out/host.txt; nproc >> gpurun_out/host.txt; lscpu | grep "Model name" >> gpurun_out/host.txt
(python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log)
(timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log)
tail -n 3 gpurun_out/smoke.log; tail -n 3 gpurun_out/pytest_gpu_full.log
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py > gpurun_out/final_s128.json 2> gpurun_out/final_s128.err
python bench.py --size 256 --steps 60 --warmup 5 > gpurun_out/final_s256.json 2> gpurun_out/final_s256.err
python bench.py --size 256 --regions 16 --balance 1 --cost 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/final_cfg3.json 2> gpurun_out/final_cfg3.err
python bench.py --size 320 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/final_s320.json 2> gpurun_out/final_s320.err
# launch list (kernel shares of the step) and one full capture per kernel; numbers printed under ncu are never bench values
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 50 --csv --log-file gpurun_out/final_launches_s128.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_force|k_kinematics|k_material|k_node" -s 40 -c 4 -f -o gpurun_out/prof_s128_final python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
ncu --set full --clock-control none -k regex:"k_force|k_kinematics|k_material|k_node" -s 16 -c 4 -f -o gpurun_out/prof_s256_final python bench.py --size 256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final256.log 2>&1
# sanitizers on a small problem, host and device-side setup
B=./lulesh_b200/bin/lulesh_b200
for tool in memcheck racecheck initcheck; do
  (echo "# compute-sanitizer --tool $tool $B -s 12 -i 12 -r 5 -c 2 [--device-setup] -q";
   timeout 280 compute-sanitizer --tool $tool $B -s 12 -i 12 -r 5 -c 2 -q 2>&1 | tail -n 4;
   timeout 280 compute-sanitizer --tool $tool $B -s 12 -i 12 -r 5 -c 2 --device-setup -q 2>&1 | tail -n 4;
   echo "$tool exit $?") > gpurun_out/sanitizer_$tool.log 2>&1
done
