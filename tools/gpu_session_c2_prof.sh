mkdir -p gpurun_out/r02z
export APM_LIB=$PWD/build_variants/libapm_normal.so
timeout 300 ncu --set full --import-source on --clock-control none -k regex:free_run_kernel -s 1 -c 1 -o gpurun_out/r02z/c2_free python tools/prof_c2.py 1 50 > gpurun_out/r02z/ncu_c2.log 2>&1; tail -2 gpurun_out/r02z/ncu_c2.log
