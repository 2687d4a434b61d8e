mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import lulesh_b200 as lb
d = lb.Device(lb.Domain(13, 16, 1, 8))
d.run(12)
print("cycles", d.scalars.cycle, d.download("e")[0])
d.set_debug(True); d.step(); d.kernel("force"); d.kernel("node", 1); d.kernel("kinematics"); d.kernel("material")
d.close()
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python /tmp/san.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 3 python /tmp/san.py > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck exit $?" >> gpurun_out/sanitizer_initcheck.log
tail -n 4 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log gpurun_out/sanitizer_initcheck.log
