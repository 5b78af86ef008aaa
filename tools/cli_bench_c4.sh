#!/bin/bash
# Config C4 end to end through the APEMoST-style executables: pulse_vrot on a synthetic spectrum
# (2000 bins on [90, 110]), N_BETA=20: calibrate_first, calibrate_rest, run
# (MAX_ITERATIONS iterations, every dump file written), analyse -> thermodynamic-integration
# evidence.  Where oracle/_ref/pulse_vrot_c4.exe exists (the UNMODIFIED reference, built in the
# container by `make -C oracle ref MODELS=pulse_vrot SUFFIX=_c4 CCFLAGS="-DN_BETA=20
# -DMAX_ITERATIONS=<n>"`) the same four phases are run with it on the box's host cores, from its
# own calibration, and the two evidence lines are printed side by side.
# Usage: tools/cli_bench_c4.sh [MAX_ITERATIONS=100000]   (GPU box)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
ITER=${1:-100000}
OUT=$(mktemp -d)
make -s -C "$ROOT/apemost_b200/host" OUT="$OUT" CCFLAGS="-DN_BETA=20 -DMAX_ITERATIONS=$ITER" "$OUT/pulse_vrot.exe"
cd "$OUT"
python - "$ROOT" <<'PY'
import sys
sys.path.insert(0, sys.argv[1] + "/tools"); sys.path.insert(0, sys.argv[1] + "/tests"); sys.path.insert(0, sys.argv[1])
import numpy as np
import small_bench
np.savetxt("data", small_bench.pulse_spectrum(2000), fmt="%.17e", delimiter="\t")
PY
printf '0.5\t0.01\t5.0\tlifetime\t-1.0\n0.0\t-10.0\t10.0\tp1\t-1.0\n0.4\t0.0\t2.0\tvrot\t-1.0\n98.0\t95.0\t100.0\tf1\t-1.0\n5.0\t0.01\t20.0\th1\t-1.0\n103.0\t100.0\t106.0\tf2\t-1.0\n3.0\t0.01\t20.0\th2\t-1.0\n' > params
run_phases() { # $1 = label, $2 = executable
	for phase in calibrate_first calibrate_rest run analyse; do
		s=$(date +%s%N)
		GSL_RNG_SEED=1 $2 $phase > $1.$phase.log 2>&1 || { echo "$1 $phase FAILED"; tail -3 $1.$phase.log; return 0; }
		e=$(date +%s%N)
		echo "$1 $phase: $(( (e - s) / 1000000 )) ms"
	done
	echo "$1 chain-steps: $((ITER * 20)); dump bytes: $(cat *.dump | wc -c)"
	grep -a -o "Model probability.*" $1.analyse.log || true
	grep -a "on-device accumulators" $1.analyse.log || true
}
run_phases gpu ./pulse_vrot.exe
# where the calibration phases' wall time goes (engine time versus CUDA start-up / shutdown)
mkdir cal_timing && cp data params cal_timing/ && (cd cal_timing && for phase in calibrate_first calibrate_rest; do
	echo "timing of $phase:"; GSL_RNG_SEED=1 APM_HOST_TIMING=1 ../pulse_vrot.exe $phase 2>&1 >/dev/null | grep timing || true; done)
GSL_RNG_SEED=1 APM_HOST_TIMING=1 ./pulse_vrot.exe run 2>&1 >/dev/null | grep timing || true
REF="$ROOT/oracle/_ref/pulse_vrot_c4.exe"
if [ -x "$REF" ]; then
	mkdir ref && cp data params ref/ && cd ref
	echo "reference on $(nproc) host cores (OpenMP):"
	run_phases ref "$REF"
	# the same ladder and start for both engines: the reference without its OpenMP loop-counter race
	# (one thread) and the GPU engine, each running from the reference's calibration
	echo "reference run from its calibration, ONE thread:"
	s=$(date +%s%N); GSL_RNG_SEED=2 OMP_NUM_THREADS=1 "$REF" run > ref1.run.log 2>&1; e=$(date +%s%N)
	echo "ref1 run: $(( (e - s) / 1000000 )) ms; dump bytes: $(cat *.dump | wc -c)"
	"$REF" analyse 2>&1 | grep -a -o "Model probability.*" || true
	mkdir ../gpu_on_ref && cp data params calibration_results ../gpu_on_ref/ && cd ../gpu_on_ref
	echo "GPU engine run from the reference's calibration:"
	GSL_RNG_SEED=3 ../pulse_vrot.exe run > run.log 2>&1
	../pulse_vrot.exe analyse 2>&1 | grep -a "Model probability" || true
	cd ..
fi
rm -rf "$OUT"
