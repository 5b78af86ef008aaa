set -x
mkdir -p gpurun_out/r02b
timeout 900 python -m pytest tests -m gpu -x -q -k "c2_phases or normal or redraws or adapt or continues or reproducible" > gpurun_out/r02b/pytest_c2.log 2>&1; tail -5 gpurun_out/r02b/pytest_c2.log
SMALL_BENCH_ONLY="C2" timeout 600 python tools/small_bench.py > gpurun_out/r02b/small_c2.log 2>&1; tail -5 gpurun_out/r02b/small_c2.log
cp gpurun_out/small_bench.json gpurun_out/r02b/small_c2.json
python tools/startup_probe.py > gpurun_out/r02b/startup.log 2>&1; cat gpurun_out/r02b/startup.log
CUDA_MODULE_LOADING=EAGER python tools/startup_probe.py > gpurun_out/r02b/startup_eager.log 2>&1; cat gpurun_out/r02b/startup_eager.log
timeout 900 bash tools/cli_bench_c4.sh 20000 > gpurun_out/r02b/cli_c4.log 2>&1; head -30 gpurun_out/r02b/cli_c4.log
ncu --set full --import-source on --clock-control none -k regex:free_run_kernel -c 1 -o gpurun_out/r02b/c2_free python tools/prof_c2.py 1 50 > gpurun_out/r02b/ncu_c2.log 2>&1; tail -3 gpurun_out/r02b/ncu_c2.log
