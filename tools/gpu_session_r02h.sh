set -x
mkdir -p gpurun_out/r02h
timeout 2400 python -m pytest tests -m gpu -x -q -s --durations=10 > gpurun_out/r02h/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h/pytest_gpu.log
grep -E "z-scores|passed|failed|rc=" gpurun_out/r02h/pytest_gpu.log | tail -8
SMALL_BENCH_ONLY="C2" timeout 600 python tools/small_bench.py > gpurun_out/r02h/small_c2.log 2>&1; tail -3 gpurun_out/r02h/small_c2.log
