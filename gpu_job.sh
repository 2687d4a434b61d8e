mkdir -p gpurun_out
B=./lulesh_b200/bin/lulesh_b200
for tool in memcheck racecheck initcheck; do
  (echo "# compute-sanitizer --tool $tool $B -s 12 -i 12 -r 5 -c 2 [--device-setup] -q (current build)";
   timeout 280 compute-sanitizer --tool $tool $B -s 12 -i 12 -r 5 -c 2 -q 2>&1 | tail -n 4;
   timeout 280 compute-sanitizer --tool $tool $B -s 12 -i 12 -r 5 -c 2 --device-setup -q 2>&1 | tail -n 4;
   echo "$tool exit $?") > gpurun_out/sanitizer_$tool.log 2>&1
  tail -n 3 gpurun_out/sanitizer_$tool.log
done
