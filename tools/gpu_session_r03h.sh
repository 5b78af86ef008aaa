mkdir -p gpurun_out/r03h
APM_LIB=build_variants/libapm_normal_t512.so SMALL_BENCH_ONLY="C2 normal" timeout 300 python tools/small_bench.py > gpurun_out/r03h/small_c2.log 2>&1; tail -n 2 gpurun_out/r03h/small_c2.log
APEMOST_GPU_LIB=$PWD/build_variants/libapm_normal_t512.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "normal or c2_phases or data_free or redraws" > gpurun_out/r03h/pytest_c2.log 2>&1; tail -n 3 gpurun_out/r03h/pytest_c2.log
