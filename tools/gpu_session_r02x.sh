set -x
mkdir -p gpurun_out/r03a
timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r03a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03a/pytest_gpu.log
grep -E "passed|failed|rc=|Error" gpurun_out/r03a/pytest_gpu.log | tail -6
SMALL_BENCH_ONLY="C2" timeout 600 python tools/small_bench.py > gpurun_out/r03a/small_c2.log 2>&1; tail -2 gpurun_out/r03a/small_c2.log
