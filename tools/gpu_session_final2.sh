mkdir -p gpurun_out/r03x
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r03x/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03x/pytest_gpu.log
tail -n 3 gpurun_out/r03x/pytest_gpu.log
