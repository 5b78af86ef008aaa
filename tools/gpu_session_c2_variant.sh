mkdir -p gpurun_out/r02z
export APM_LIB=$PWD/build_variants/libapm_normal.so
SMALL_BENCH_ONLY="C2" timeout 300 python tools/small_bench.py > gpurun_out/r02z/small_c2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02z/small_c2.log
