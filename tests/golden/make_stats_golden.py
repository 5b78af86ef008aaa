#!/usr/bin/env python
"""Full-configuration statistical fixtures from the UNMODIFIED reference (oracle/_ref; needs
/root/reference): the second half of "correctness" in BASELINE.json's north_star -- posterior
means / variances and the thermodynamic-integration evidence of the reference's own engine, with
their seed-to-seed scatter, at the configurations' real sizes:

  c1_stats.json   C1: simplesin on tests/testlc.dat, N_BETA = 20, default burn-in and calibration,
                  MAX_ITERATIONS = 20000
  c4_stats.json   C4: pulse_vrot on the 2000-bin synthetic spectrum (tools/small_bench.pulse_spectrum,
                  seed 4242), N_BETA = 20, MAX_ITERATIONS = 100000

Protocol: ONE calibration (calibrate_first + calibrate_rest, GSL_RNG_SEED=1) fixes the ladder, the
step widths and the start points; then `run` + `analyse` are repeated for N_SEEDS values of
GSL_RNG_SEED from that same calibration_results, with OMP_NUM_THREADS=1 (the OpenMP build's shared
loop counter skips steps).  Recorded per seed: ln Z as `analyse` prints it, and the
mean and variance of every parameter of chain 0 (beta = 1) over the whole run, computed from the
reference's <name>-chain-0.prob.dump files.  The GPU tests run the engine from the same calibration
over independent ensembles and compare the means within 4 standard errors of the difference.

Usage: python tests/golden/make_stats_golden.py [c1] [c4]      (~10 minutes on 8 cores)
"""
import concurrent.futures
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle_binding import ref_exe, write_params_file, write_data_file  # noqa: E402
import make_golden  # noqa: E402

N_SEEDS = 24


def one_seed(exe, base, seed, names):
    d = tempfile.mkdtemp(prefix="apm_stat_")
    try:
        for f in ("params", "data", "calibration_results"):
            shutil.copy(os.path.join(base, f), os.path.join(d, f))
        env = dict(os.environ, GSL_RNG_SEED=str(seed), OMP_NUM_THREADS="1")
        subprocess.run([exe, "run"], cwd=d, env=env, check=True, capture_output=True)
        r = subprocess.run([exe, "analyse"], cwd=d, env=env, check=True, capture_output=True, text=True)
        m = re.search(r"Model probability ln\(p\(D\|M, I\)\): \[about 10\^(-?\d+)\] (-?[\d.]+)", r.stdout)
        mean, var = [], []
        for name in names:
            v = np.loadtxt(os.path.join(d, "%s-chain-0.prob.dump" % name))
            mean.append(float(v.mean()))
            var.append(float(v.var()))
        return dict(seed=seed, lnz=float(m.group(2)), n=int(len(v)), mean=mean, var=var)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def stats_fixture(model, rows, data, data_file, n_beta, iters, suffix):
    flags = f"-DN_BETA={n_beta} -DMAX_ITERATIONS={iters}"
    exe = ref_exe(model, ccflags=flags, suffix=suffix)
    names = [r[3] for r in rows]
    base = tempfile.mkdtemp(prefix="apm_stat_base_")
    try:
        write_params_file(os.path.join(base, "params"), rows)
        if data_file:
            shutil.copy(os.path.join(HERE, data_file), os.path.join(base, "data"))
        else:
            write_data_file(os.path.join(base, "data"), data)
        env = dict(os.environ, GSL_RNG_SEED="1", OMP_NUM_THREADS="1")
        for phase in ("calibrate_first", "calibrate_rest"):
            subprocess.run([exe, phase], cwd=base, env=env, check=True, capture_output=True)
        cal = open(os.path.join(base, "calibration_results")).read()
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(N_SEEDS, os.cpu_count() or 1)) as ex:
            runs = list(ex.map(lambda s: one_seed(exe, base, s, names), range(1, N_SEEDS + 1)))
    finally:
        shutil.rmtree(base, ignore_errors=True)
    lnz = np.array([r["lnz"] for r in runs])
    mean = np.array([r["mean"] for r in runs])
    var = np.array([r["var"] for r in runs])
    return dict(
        model=model, rows=[list(r) for r in rows], data_file=data_file,
        data=None if data_file else make_golden.frepr(data), n_cols=2,
        config=dict(N_BETA=n_beta, MAX_ITERATIONS=iters, calibration_seed=1, threads=1), ccflags=flags,
        calibration_results=cal, runs=runs,
        summary=dict(n_seeds=len(runs), lnz_mean=float(lnz.mean()), lnz_sd=float(lnz.std(ddof=1)),
                     param_mean=mean.mean(axis=0).tolist(), param_mean_sd=mean.std(axis=0, ddof=1).tolist(),
                     param_var=var.mean(axis=0).tolist(), param_var_sd=var.std(axis=0, ddof=1).tolist()),
        source=f"oracle/_ref/{model}{suffix}.exe = unmodified reference, {flags}; this script")


def main():
    which = sys.argv[1:] or ["c1", "c4"]
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "_build", "gsl_compat.o")],
                   check=True)
    if "c1" in which:
        fx = stats_fixture("simplesin", make_golden.C1_ROWS, None, "testlc.dat", 20, 20000, "_stat1")
        json.dump(fx, open(os.path.join(HERE, "c1_stats.json"), "w"), indent=1)
        print("wrote c1_stats.json", fx["summary"])
    if "c4" in which:
        import small_bench
        fx = stats_fixture("pulse_vrot", make_golden.MODELS["pulse_vrot"]["rows"], small_bench.pulse_spectrum(2000), None,
                           20, 100000, "_stat4")
        json.dump(fx, open(os.path.join(HERE, "c4_stats.json"), "w"), indent=1)
        print("wrote c4_stats.json", fx["summary"])


if __name__ == "__main__":
    main()
