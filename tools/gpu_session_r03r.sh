mkdir -p gpurun_out/r03r
for s in 13 20 26 37 52 74; do echo "APM_SPLITS=$s"; APM_SPLITS=$s timeout 200 python tools/prof_eval.py 2>&1 | tail -n 1; done > gpurun_out/r03r/splits.log 2>&1; cat gpurun_out/r03r/splits.log
