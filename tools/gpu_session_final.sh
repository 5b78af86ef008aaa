mkdir -p gpurun_out/r03t
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r03t/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03t/pytest_gpu.log
tail -n 3 gpurun_out/r03t/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03t/smoke.log 2>&1; tail -n 1 gpurun_out/r03t/smoke.log
timeout 900 python bench.py > gpurun_out/r03t/bench_n1.json 2> gpurun_out/r03t/bench_n1.err; tail -c 300 gpurun_out/r03t/bench_n1.json
SMALL_BENCH_ONLY="C2" timeout 300 python tools/small_bench.py > gpurun_out/r03t/small_c2.log 2>&1; tail -n 2 gpurun_out/r03t/small_c2.log
