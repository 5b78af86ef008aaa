set -x
mkdir -p gpurun_out/r02p
timeout 1500 python -m pytest tests -m gpu -x -q -k "calibration or steps or group or phases_on_gpu or alternate" > gpurun_out/r02p/pytest_cal.log 2>&1; tail -3 gpurun_out/r02p/pytest_cal.log
python tools/prof_eval.py - 4 > gpurun_out/r02p/prof_eval.log 2>&1; tail -3 gpurun_out/r02p/prof_eval.log
timeout 900 bash tools/cli_bench_c4.sh 20000 > gpurun_out/r02p/cli_c4.log 2>&1; grep -E "calibrat|gpu " gpurun_out/r02p/cli_c4.log | head -20
