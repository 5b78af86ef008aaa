set -x
mkdir -p gpurun_out/r02n
timeout 2400 python -m pytest tests -m gpu -x -q -s --durations=10 > gpurun_out/r02n/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n/pytest_gpu.log
grep -E "z-scores|passed|failed|rc=|Error|error" gpurun_out/r02n/pytest_gpu.log | tail -12
timeout 900 bash tools/cli_bench_c4.sh > gpurun_out/r02n/cli_c4.log 2>&1; head -40 gpurun_out/r02n/cli_c4.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n/smoke.log 2>&1; tail -2 gpurun_out/r02n/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02n/bench_n1.json 2> gpurun_out/r02n/bench_n1.err; tail -c 2500 gpurun_out/r02n/bench_n1.json
