mkdir -p gpurun_out/r03n
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "error_codes" > gpurun_out/r03n/pytest_err.log 2>&1; tail -n 25 gpurun_out/r03n/pytest_err.log | cut -c1-200
