mkdir -p gpurun_out/r03m
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r03m/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03m/pytest_gpu.log
tail -n 3 gpurun_out/r03m/pytest_gpu.log
SMALL_BENCH_ONLY="C2" timeout 300 python tools/small_bench.py > gpurun_out/r03m/small_c2.log 2>&1; tail -n 2 gpurun_out/r03m/small_c2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
