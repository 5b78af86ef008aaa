set -x
mkdir -p gpurun_out/r02r
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02r/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r/pytest_gpu.log
grep -E "passed|failed|rc=|Error" gpurun_out/r02r/pytest_gpu.log | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02r/bench_n1.json 2> gpurun_out/r02r/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r/bench_n1.json').read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"],"kernel ms",d["roofline"]["kernel_ms_per_launch"],"share",d["roofline"]["kernel_share_of_step"])
PY
python tools/prof_eval.py - 3 pulse_vrot > gpurun_out/r02r/prof_pulse.log 2>&1; tail -2 gpurun_out/r02r/prof_pulse.log
