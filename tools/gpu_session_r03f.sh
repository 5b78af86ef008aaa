set -x
mkdir -p gpurun_out/r03f
timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r03f/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03f/pytest_gpu.log
grep -E "passed|failed|rc=|Error" gpurun_out/r03f/pytest_gpu.log | tail -n 6
timeout 900 python tools/small_bench.py > gpurun_out/r03f/small_all.log 2>&1; tail -n 12 gpurun_out/r03f/small_all.log; cp gpurun_out/small_bench.json gpurun_out/r03f/small_bench.json
timeout 300 python tools/sanitize_small.py > gpurun_out/r03f/sanitize.log 2>&1; tail -n 6 gpurun_out/r03f/sanitize.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03f/smoke.log 2>&1; tail -n 2 gpurun_out/r03f/smoke.log
timeout 900 python bench.py > gpurun_out/r03f/bench_n1.json 2> gpurun_out/r03f/bench_n1.err; tail -c 1500 gpurun_out/r03f/bench_n1.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:free_run_kernel -s 1 -c 1 -o gpurun_out/r03f/c2_free python tools/prof_c2.py 1 50 > gpurun_out/r03f/ncu_c2.log 2>&1; tail -n 2 gpurun_out/r03f/ncu_c2.log
