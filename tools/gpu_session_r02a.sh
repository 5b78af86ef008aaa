set -x
mkdir -p gpurun_out/r02a
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest_gpu.log
tail -30 gpurun_out/r02a/pytest_gpu.log
SMALL_BENCH_ONLY="C2" timeout 600 python tools/small_bench.py > gpurun_out/r02a/small_c2.log 2>&1; tail -5 gpurun_out/r02a/small_c2.log
cp gpurun_out/small_bench.json gpurun_out/r02a/small_c2.json
timeout 900 bash tools/cli_bench_c4.sh > gpurun_out/r02a/cli_c4.log 2>&1; head -12 gpurun_out/r02a/cli_c4.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a/bench_n1.json 2> gpurun_out/r02a/bench_n1.err; tail -c 3000 gpurun_out/r02a/bench_n1.json
