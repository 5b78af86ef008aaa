mkdir -p gpurun_out/r02k
for s in 13 8 4 2 1; do echo "APM_SPLITS=$s"; APM_SPLITS=$s python tools/prof_eval.py - 4 2>&1 | tail -2; done > gpurun_out/r02k/splits_sweep.log 2>&1
cat gpurun_out/r02k/splits_sweep.log
