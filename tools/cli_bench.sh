#!/bin/bash
# End-to-end timing of the APEMoST-style executable on config C1 (simplesin on the reference's
# light curve, N_BETA=20): calibrate_first, calibrate_rest, run (MAX_ITERATIONS iterations, every
# dump file written), analyse.  Usage: tools/cli_bench.sh [MAX_ITERATIONS=200000]   (GPU box)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
ITER=${1:-200000}
OUT=$(mktemp -d)
make -s -C "$ROOT/apemost_b200/host" OUT="$OUT" CCFLAGS="-DN_BETA=20 -DMAX_ITERATIONS=$ITER" "$OUT/simplesin.exe"
cd "$OUT"
cp "$ROOT/tests/golden/testlc.dat" data
printf '1.0\t0.0\t3.0\tamplitude\t-1.0\n15.2\t4.0\t24.0\tfrequency\t0.001\n0.25\t0.0\t1.0\tphase\t-1.0\n0.0\t-1.0\t1.0\toffset\t-1.0\n' > params
for phase in calibrate_first calibrate_rest run analyse; do
	s=$(date +%s%N)
	GSL_RNG_SEED=1 ./simplesin.exe $phase > $phase.log 2>&1
	e=$(date +%s%N)
	echo "$phase: $(( (e - s) / 1000000 )) ms"
done
s=$(date +%s%N); GSL_RNG_SEED=1 OMP_NUM_THREADS=1 ./simplesin.exe run > run1.log 2>&1; e=$(date +%s%N)
echo "run (OMP_NUM_THREADS=1: dump files formatted by one thread): $(( (e - s) / 1000000 )) ms"
GSL_RNG_SEED=1 APM_HOST_TIMING=1 ./simplesin.exe calibrate_rest 2>&1 >/dev/null | grep timing || true
GSL_RNG_SEED=1 APM_HOST_TIMING=1 ./simplesin.exe run 2>&1 >/dev/null | grep timing || true
echo "chain-steps: $((ITER * 20)); dump bytes: $(cat *.dump | wc -c)"
grep -a -o "Model probability.*" analyse.log || true
rm -rf "$OUT"
