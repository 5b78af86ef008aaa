mkdir -p gpurun_out/r03e
export APM_LIB=$PWD/build_variants/libapm_normal_t512.so
timeout 300 ncu --set full --import-source on --clock-control none -k regex:free_run_kernel -s 1 -c 1 -o gpurun_out/r03e/c2_free python tools/prof_c2.py 1 50 > gpurun_out/r03e/ncu_c2.log 2>&1; tail -n 2 gpurun_out/r03e/ncu_c2.log
