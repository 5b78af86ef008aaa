mkdir -p gpurun_out/r03h
APM_LIB=build_variants/libapm_normal_clk.so timeout 120 python tools/prof_c2.py 1 200 > gpurun_out/r03h/clk.log 2>&1; echo "rc=$?"; tail -n 20 gpurun_out/r03h/clk.log | sort | cut -c1-150
APM_LIB=build_variants/libapm_normal_t512.so SMALL_BENCH_ONLY="C2 normal" timeout 200 python tools/small_bench.py > gpurun_out/r03h/small_c2.log 2>&1; tail -n 2 gpurun_out/r03h/small_c2.log
APEMOST_GPU_LIB=$PWD/build_variants/libapm_normal_t512.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "normal or c2_phases or data_free or redraws" > gpurun_out/r03h/pytest_c2.log 2>&1; tail -n 3 gpurun_out/r03h/pytest_c2.log
