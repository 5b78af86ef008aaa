mkdir -p gpurun_out/r03p
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r03p/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03p/pytest_gpu.log
tail -n 3 gpurun_out/r03p/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r03p/bench_n1.json 2> gpurun_out/r03p/bench_n1.err; tail -c 300 gpurun_out/r03p/bench_n1.json
timeout 300 python tools/prof_eval.py > gpurun_out/r03p/prof_eval.log 2>&1; tail -n 2 gpurun_out/r03p/prof_eval.log
