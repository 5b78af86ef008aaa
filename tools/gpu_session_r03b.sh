set -x
mkdir -p gpurun_out/r03b
for a in reg; do
APM_LIB=build_variants/libapm_normal_$a.so SMALL_BENCH_ONLY="C2" timeout 300 python tools/small_bench.py > gpurun_out/r03b/small_c2_$a.log 2>&1; tail -n 2 gpurun_out/r03b/small_c2_$a.log
done
timeout 900 python -m pytest tests -m gpu -x -q -k "normal or c2 or free or redraw or marginal or data_free" > gpurun_out/r03b/pytest_c2.log 2>&1; tail -n 3 gpurun_out/r03b/pytest_c2.log
timeout 200 python tools/sanitize_small.py 5 2>&1 | tail -n 2
