set -x
mkdir -p gpurun_out/r02v
export APM_LIB=$PWD/build_variants/libapm_normal.so
SMALL_BENCH_ONLY="C2" timeout 600 python tools/small_bench.py > gpurun_out/r02v/small_c2.log 2>&1; tail -5 gpurun_out/r02v/small_c2.log
ncu --set full --import-source on --clock-control none -k regex:free_run_kernel -s 1 -c 1 -o gpurun_out/r02v/c2_free python tools/prof_c2.py 1 50 > gpurun_out/r02v/ncu_c2.log 2>&1; tail -3 gpurun_out/r02v/ncu_c2.log
