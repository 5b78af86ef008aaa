mkdir -p gpurun_out/r02w
for att in 3 4; do echo "ATT=$att"; APM_LIB=$PWD/build_variants/libapm_normal_att$att.so SMALL_BENCH_ONLY="C2" python tools/small_bench.py 2>&1 | tail -2; done > gpurun_out/r02w/c2_att.log 2>&1
cat gpurun_out/r02w/c2_att.log
