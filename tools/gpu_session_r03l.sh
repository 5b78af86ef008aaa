mkdir -p gpurun_out/r03l
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r03l/bench_plain.json 2>&1; tail -c 200 gpurun_out/r03l/bench_plain.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03l/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r03l/ncu_bench.log 2>&1; tail -c 200 gpurun_out/r03l/ncu_bench.log
timeout 300 python tools/prof_eval.py > gpurun_out/r03l/prof_eval.log 2>&1; tail -n 2 gpurun_out/r03l/prof_eval.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:loglik_tiled -c 1 -o gpurun_out/r03l/loglik_full python tools/prof_eval.py > gpurun_out/r03l/ncu_loglik.log 2>&1; tail -n 2 gpurun_out/r03l/ncu_loglik.log
