mkdir -p gpurun_out/r02l
for s in 13 2; do echo "APM_SPLITS=$s"; APM_SPLITS=$s python tools/prof_eval.py build_variants/libapm_ss5.so 4 2>&1 | tail -2; done > gpurun_out/r02l/splits_sweep.log 2>&1
cat gpurun_out/r02l/splits_sweep.log
