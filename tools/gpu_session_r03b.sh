mkdir -p gpurun_out/r03b
APM_LIB=build_variants/libapm_normal_clk.so timeout 300 python tools/prof_c2.py 1 200 > gpurun_out/r03b/clk.log 2>&1; tail -n 20 gpurun_out/r03b/clk.log | sort | cut -c1-150
for a in t512; do
APM_LIB=build_variants/libapm_normal_$a.so SMALL_BENCH_ONLY="C2 normal" timeout 300 python tools/small_bench.py > gpurun_out/r03b/small_c2_$a.log 2>&1; echo $a; tail -n 2 gpurun_out/r03b/small_c2_$a.log
done
