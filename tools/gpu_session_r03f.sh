set -x
mkdir -p gpurun_out/r03f
timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r03f/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03f/pytest_gpu.log
grep -E "passed|failed|rc=|Error" gpurun_out/r03f/pytest_gpu.log | tail -n 6
SMALL_BENCH_ONLY="C2" timeout 600 python tools/small_bench.py > gpurun_out/r03f/small_c2.log 2>&1; tail -n 2 gpurun_out/r03f/small_c2.log
timeout 300 python tools/sanitize_small.py 5 > gpurun_out/r03f/sanitize5.log 2>&1; tail -n 2 gpurun_out/r03f/sanitize5.log
