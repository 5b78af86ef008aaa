mkdir -p gpurun_out/r03s
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03s/smoke.log 2>&1; tail -n 1 gpurun_out/r03s/smoke.log
timeout 900 bash tools/cli_bench_c4.sh > gpurun_out/r03s/cli_c4.log 2>&1; tail -n 25 gpurun_out/r03s/cli_c4.log | cut -c1-200
