mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/pytest_gpu_n1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_n1.log)
tail -n 5 gpurun_out/pytest_gpu_n1.log
( time ./lulesh_b200/bin/lulesh_b200 -s 256 -i 20 --device-setup -q ) 2>&1 | tail -3
( time ./lulesh_b200/bin/lulesh_b200 -s 256 -i 20 -q ) 2>&1 | tail -3
