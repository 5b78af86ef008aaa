set -x
mkdir -p gpurun_out/r02j
build_variants/fp64_peak_probe > gpurun_out/r02j/fp64_peak_probe.json 2> gpurun_out/r02j/fp64_peak_probe.err; cat gpurun_out/r02j/fp64_peak_probe.json
timeout 1200 python -m pytest tests/test_gpu_host.py -m gpu -x -q -k "alternate" > gpurun_out/r02j/pytest_altcal.log 2>&1; tail -3 gpurun_out/r02j/pytest_altcal.log
python tools/prof_eval.py > gpurun_out/r02j/prof_eval.log 2>&1; tail -2 gpurun_out/r02j/prof_eval.log
ncu --set full --import-source on --clock-control none -k regex:loglik_tiled -c 1 -o gpurun_out/r02j/loglik_full python tools/prof_eval.py > gpurun_out/r02j/ncu_loglik.log 2>&1; tail -2 gpurun_out/r02j/ncu_loglik.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02j/ncu_bench.log 2>&1; tail -c 300 gpurun_out/r02j/ncu_bench.log
