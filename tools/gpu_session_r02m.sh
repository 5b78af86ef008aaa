mkdir -p gpurun_out/r02m
ncu --set full --import-source on --clock-control none -k regex:loglik_tiled -c 1 -o gpurun_out/r02m/loglik_ws python tools/prof_eval.py build_variants/libapm_ss5.so 2 > gpurun_out/r02m/ncu.log 2>&1; tail -2 gpurun_out/r02m/ncu.log
