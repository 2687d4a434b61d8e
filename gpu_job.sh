mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_n1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_n1.log)
tail -n 6 gpurun_out/pytest_gpu_n1.log
