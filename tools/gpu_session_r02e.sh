set -x
mkdir -p gpurun_out/r02e
timeout 900 python -m pytest tests -m gpu -x -q -k "c2_phases or normal or redraws or adapt or continues or reproducible" > gpurun_out/r02e/pytest_c2.log 2>&1; tail -5 gpurun_out/r02e/pytest_c2.log
ncu --set full --import-source on --clock-control none -k regex:free_run_kernel -s 1 -c 1 -o gpurun_out/r02e/c2_free python tools/prof_c2.py 1 50 > gpurun_out/r02e/ncu_c2.log 2>&1; tail -3 gpurun_out/r02e/ncu_c2.log
