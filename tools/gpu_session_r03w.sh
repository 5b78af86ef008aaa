mkdir -p gpurun_out/r03w
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_run" --durations=3 > gpurun_out/r03w/pytest_long.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/r03w/pytest_long.log | cut -c1-220
