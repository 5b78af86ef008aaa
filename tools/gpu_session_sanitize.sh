set -x
mkdir -p gpurun_out/r02u
python tools/sanitize_small.py > gpurun_out/r02u/plain.log 2>&1; tail -6 gpurun_out/r02u/plain.log
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r02u/memcheck.log 2>&1; tail -4 gpurun_out/r02u/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2 3 5 > gpurun_out/r02u/racecheck.log 2>&1; tail -4 gpurun_out/r02u/racecheck.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_small.py 1 3 5 > gpurun_out/r02u/synccheck.log 2>&1; tail -4 gpurun_out/r02u/synccheck.log
