APM_LIB=$PWD/build_variants/libapm_normal_k8.so SMALL_BENCH_ONLY="C2 normal 1x64" timeout 300 python tools/small_bench.py 2>&1 | tail -1
