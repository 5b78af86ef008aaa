"""The C host layer (apemost_b200/host) pinned against the unmodified reference, without a GPU.

The host layer's sources are compiled UNCHANGED and linked, instead of libapemost_gpu.so, against
tests/host_shim/apm_gpu_over_oracle.c, which serves the same C ABI from the CPU oracle in
MT19937 mode (the oracle itself is byte-pinned to the reference by test_oracle_golden.py).  Then
`<model>.exe calibrate_first / calibrate_rest / run / analyse` must leave exactly the files the
reference left (tests/golden/*_phases.json, recorded by tests/golden/make_golden.py):
calibration_results after every phase, every dump, params_suggested, calibration_summary,
calibration_progress.data, acceptance_rate.dump[.gnuplot], the histograms, the gnuplot script and
the evidence line.  On a GPU box tests/test_gpu_host.py runs the same executables over the real
engine."""
import hashlib
import json
import os
import re
import subprocess

import pytest

from oracle_binding import build_oracle, write_data_file, write_params_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
HOST = os.path.join(ROOT, "apemost_b200", "host")
BUILD = os.path.join(ROOT, "oracle", "_build")
FIXTURES = ["c1_phases", "c1_circular_phases", "c1_logistic_phases", "c1_uniform_phases", "c4_phases", "c2_phases",
            "c1_adapt_phases", "c1_randomswap_phases", "c1_altcal_phases", "c1_multilin_phases", "c1_quadratic_phases"]
MODEL_IDS = {"simplesin": 0, "simplesin5": 1, "normal": 2, "pulse_vrot": 3, "simplesin2": 4, "pulse": 5}


def build_host_over_oracle(name, model, ccflags):
    """gcc: host layer + shim + liboracle.so -> oracle/_build/host_<name>.exe"""
    build_oracle()
    exe = os.path.join(BUILD, f"host_{name}.exe")
    src = [os.path.join(HOST, f) for f in ("apm_main.c", "apm_chainobj.c", "apm_files.c", "apm_phases.c",
                                           "apm_analyse.c", "apm_assess.c", "apm_calibrate_alt.c", "apm_calibrate_multilin.c",
                                           "apm_calibrate_quadratic.c", "apm_fastfmt.c")]
    src += [os.path.join(ROOT, "apemost_b200", "compat", "gsl", "gsl_compat.c"),
            os.path.join(ROOT, "tests", "host_shim", "apm_gpu_over_oracle.c")]
    cmd = ["gcc", "-O2", "-std=gnu99", "-fopenmp", "-pthread", "-I", os.path.join(HOST, "include"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "apemost_b200", "compat"), "-I", os.path.join(ROOT, "oracle"),
           f"-DAPM_MODEL_ID={MODEL_IDS[model]}", f'-DAPM_MODEL_NAME="{model}"', *ccflags, *src,
           os.path.join(BUILD, "liboracle.so"), f"-Wl,-rpath,{BUILD}", "-lm", "-lgomp", "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("name", FIXTURES)
def test_host_layer_files_byte_identical_to_reference(name, tmp_path):
    fx = json.load(open(os.path.join(GOLDEN, name + ".json")))
    cfg = fx["config"]
    flags = [f"-D{k}={v}" for k, v in cfg.items() if k != "GSL_RNG_SEED"] + fx["ccflags_extra"].split()
    exe = build_host_over_oracle(name, fx["model"], flags)
    wd = str(tmp_path)
    write_params_file(os.path.join(wd, "params"), [tuple(r) for r in fx["rows"]])
    if fx["data_file"]:
        open(os.path.join(wd, "data"), "wb").write(open(os.path.join(GOLDEN, fx["data_file"]), "rb").read())
    else:
        import numpy as np
        write_data_file(os.path.join(wd, "data"), np.array(fx["data"], dtype=float).reshape(-1, fx["n_cols"]))
    env = dict(os.environ, GSL_RNG_SEED=str(cfg["GSL_RNG_SEED"]), APM_TEST_ORACLE_RNG="mt19937")
    if "-DRANDOMSWAP" in fx["ccflags_extra"]:
        env["APM_TEST_RANDOMSWAP"] = "1"
    for phase in ("calibrate_first", "calibrate_rest", "run"):
        subprocess.run([exe, phase], cwd=wd, env=env, check=True, capture_output=True)
        assert open(os.path.join(wd, "calibration_results")).read() == fx["phases"][phase], phase
    for fname, want in fx["dumps"].items():
        lines = open(os.path.join(wd, fname)).read().splitlines()
        assert len(lines) == want["n_lines"], fname
        assert lines[:5] == want["head"] and lines[-5:] == want["tail"], fname
        assert sha(os.path.join(wd, fname)) == want["sha256"], fname
    r = subprocess.run([exe, "analyse"], cwd=wd, env=env, check=True, capture_output=True, text=True)
    for fname, want in fx["files"].items():
        path = os.path.join(wd, fname)
        assert os.path.exists(path), fname
        if want["text"] is not None:
            assert open(path).read() == want["text"], fname
        assert sha(path) == want["sha256"], fname
    # what analyse prints: the evidence block and the per-parameter error estimates
    m = re.search(r"Model probability ln\(p\(D\|M, I\)\): \[about 10\^(-?\d+)\] (-?[\d.]+)", r.stdout)
    assert m and m.group(0) == fx["evidence_line"]
    want_block = fx["analyse_stdout"][fx["analyse_stdout"].index("Model probability"):]
    got = r.stdout[r.stdout.index("Model probability"):]
    assert got.startswith(want_block)
    assert re.findall(r"mcmc error estimate of .*", r.stdout) == re.findall(r"mcmc error estimate of .*",
                                                                           fx["analyse_stdout"])
    # new: the evidence from the accumulators of `run` agrees with the one from the 7-digit dumps
    m2 = re.search(r"on-device accumulators \(full precision\): (-?[\d.]+)", r.stdout)
    assert m2 and abs(float(m2.group(1)) - float(fx["evidence"])) < 2e-4 * max(1.0, abs(float(fx["evidence"])))


def check_analyse_from_accumulators(exe, wd, env, n_par_names):
    """after calibrate_first / calibrate_rest in `wd`: `run` with the text dumps, `analyse` from them
    (the reference's way); then the dumps are moved away and `analyse` runs again, from run_marginals
    and run_statistics (the on-device accumulators): the same <name>.histogram files
    byte for byte, the same error estimates; then a `run` with APM_NO_DUMPS=1 leaves the same
    accumulator files and no dump at all"""
    import glob
    import hashlib
    import shutil
    subprocess.run([exe, "run"], cwd=wd, env=env, check=True, capture_output=True)
    assert os.path.exists(os.path.join(wd, "run_marginals"))
    r1 = subprocess.run([exe, "analyse"], cwd=wd, env=env, check=True, capture_output=True, text=True)
    hist1 = {n: open(os.path.join(wd, n + ".histogram")).read() for n in n_par_names}
    est1 = re.findall(r"mcmc error estimate of .*", r1.stdout)
    ev1 = float(re.search(r"Model probability ln\(p\(D\|M, I\)\): \[about 10\^(-?\d+)\] (-?[\d.]+)", r1.stdout).group(2))
    marg1 = open(os.path.join(wd, "run_marginals")).read()
    stash = os.path.join(wd, "dumps_moved_away")
    os.makedirs(stash)
    dumps = glob.glob(os.path.join(wd, "*.dump"))
    assert len(dumps) > len(n_par_names)
    for f in dumps:
        shutil.move(f, stash)
    for n in n_par_names:
        os.remove(os.path.join(wd, n + ".histogram"))
    r2 = subprocess.run([exe, "analyse"], cwd=wd, env=env, check=True, capture_output=True, text=True)
    for n in n_par_names:
        assert open(os.path.join(wd, n + ".histogram")).read() == hist1[n], n
    assert re.findall(r"mcmc error estimate of .*", r2.stdout) == est1 and len(est1) == len(n_par_names)
    ev2 = float(re.search(r"Model probability ln\(p\(D\|M, I\)\): \[about 10\^(-?\d+)\] (-?[\d.]+)", r2.stdout).group(2))
    assert abs(ev2 - ev1) < 2e-4 * max(1.0, abs(ev1))   # 7-digit text against full-precision sums
    # the same run without any text dump
    os.remove(os.path.join(wd, "run_marginals"))
    subprocess.run([exe, "run"], cwd=wd, env=dict(env, APM_NO_DUMPS="1"), check=True, capture_output=True)
    assert not [f for f in glob.glob(os.path.join(wd, "*.dump")) if "acceptance_rate" not in f]
    assert open(os.path.join(wd, "run_marginals")).read() == marg1
    r3 = subprocess.run([exe, "analyse"], cwd=wd, env=env, check=True, capture_output=True, text=True)
    for n in n_par_names:
        assert open(os.path.join(wd, n + ".histogram")).read() == hist1[n], n
    assert re.findall(r"mcmc error estimate of .*", r3.stdout) == est1


@pytest.mark.parametrize("name", ["c1_phases", "c4_phases"])
def test_analyse_from_accumulators_equals_analyse_from_dumps(name, tmp_path):
    """marginal histograms, batch-means error estimates and the evidence from the
    engine's accumulators instead of the text dumps (here: the host layer over the oracle shim)"""
    import numpy as np
    fx = json.load(open(os.path.join(GOLDEN, name + ".json")))
    cfg = fx["config"]
    flags = [f"-D{k}={v}" for k, v in cfg.items() if k != "GSL_RNG_SEED"] + fx["ccflags_extra"].split()
    exe = build_host_over_oracle(name, fx["model"], flags)
    wd = str(tmp_path)
    write_params_file(os.path.join(wd, "params"), [tuple(r) for r in fx["rows"]])
    if fx["data_file"]:
        open(os.path.join(wd, "data"), "wb").write(open(os.path.join(GOLDEN, fx["data_file"]), "rb").read())
    else:
        write_data_file(os.path.join(wd, "data"), np.array(fx["data"], dtype=float).reshape(-1, fx["n_cols"]))
    env = dict(os.environ, GSL_RNG_SEED=str(cfg["GSL_RNG_SEED"]), APM_TEST_ORACLE_RNG="mt19937")
    for phase in ("calibrate_first", "calibrate_rest"):
        subprocess.run([exe, phase], cwd=wd, env=env, check=True, capture_output=True)
    check_analyse_from_accumulators(exe, wd, env, [r[3] for r in fx["rows"]])
    # and the histograms are the reference's own
    for fname, want in fx["files"].items():
        if fname.endswith(".histogram"):
            assert sha(os.path.join(wd, fname)) == want["sha256"], fname


def test_fast_e6_formatter_writes_printf_bytes(tmp_path):
    """apm_fastfmt.c ("%6e" for prob-chain<k>.dump and "%.15e" for the parameter dumps without printf) against snprintf on millions of
    values: random bit patterns, typical log-likelihoods, scaled integers, values sitting on
    rounding ties (which it must decline), powers of ten +- a few ulps"""
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
int apm_format_e6(double v, char * out);
int apm_format_e15(double v, char * out);
static uint64_t s = 88172645463325252ULL;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(void) {
	long n = 6000000, i, declined = 0, bad = 0, ties_taken = 0;
	char a[64], b[64];
	for (i = 0; i < n; i++) {
		double v;
		uint64_t r = rnd();
		switch (i % 6) {
		case 0: { union { uint64_t u; double d; } x; x.u = r; v = x.d; break; }
		case 1: v = -(double) (r % 2000000000) / 1e3 - 700000.0; break;
		case 2: v = ldexp((double) (r >> 11), (int) (rnd() % 200) - 150); break;
		case 3: { long k = (long) (r % 9000000) + 1000000; int e = (int) (rnd() % 60) - 30; v = (k + 0.5) * pow(10, e - 6); break; }
		case 4: v = pow(10, (int) (r % 40) - 20) * (1 + ((double) (int) (rnd() % 7) - 3) * 1.1102230246251565e-16); break;
		default: v = ((double) (r % 10000000) - 5e6) * 1e-7; break;
		}
		int la = apm_format_e6(v, a);
		if (la == 0) { declined++; continue; }
		if (i % 6 == 3) ties_taken++;
		a[la] = 0;
		snprintf(b, sizeof(b), "%6e", v);
		if (strcmp(a, b) != 0 && bad++ < 10) printf("MISMATCH %.17g: fast '%s' printf '%s'\n", v, a, b);
	}
	for (i = 0; i < n; i++) { /* "%.15e": the parameter dumps */
		double v;
		uint64_t r = rnd();
		switch (i % 5) {
		case 0: { union { uint64_t u; double d; } x; x.u = r; v = x.d; break; }
		case 1: v = 15.2 + ((double) (r >> 11) / 9007199254740992.0 - 0.5) * 1e-3; break;
		case 2: v = ldexp((double) (r >> 11), (int) (rnd() % 120) - 90); break;
		case 3: v = pow(10, (int) (r % 40) - 12) * (1 + ((double) (int) (rnd() % 7) - 3) * 1.1102230246251565e-16); break;
		default: v = ((double) (r % 1000000007) - 5e8) * 1e-9; break;
		}
		int la = apm_format_e15(v, a);
		if (la == 0) continue;
		a[la] = 0;
		snprintf(b, sizeof(b), "%.15e", v);
		if (strcmp(a, b) != 0 && bad++ < 10) printf("MISMATCH %.17g: fast '%s' printf '%s'\n", v, a, b);
	}
	printf("%ld values, %ld declined, %ld mismatches\n", n, declined, bad);
	return bad != 0 || declined > n / 2;
}
''')
    exe = str(tmp_path / "t.exe")
    subprocess.run(["gcc", "-O2", str(src), os.path.join(HOST, "apm_fastfmt.c"), "-lm", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout


def test_host_layer_rejects_bad_params_file(tmp_path):
    """the parser's checks (reference src/mcmc_parser.c:61-82): start outside [min, max] -> exit(1)"""
    exe = build_host_over_oracle("c1_phases", "simplesin", ["-DN_BETA=4"])
    wd = str(tmp_path)
    write_params_file(os.path.join(wd, "params"), [(5.0, 0.0, 3.0, "amplitude", -1.0)])
    open(os.path.join(wd, "data"), "w").write("0 0\n1 1\n")
    r = subprocess.run([exe, "calibrate_first"], cwd=wd, capture_output=True, text=True)
    assert r.returncode == 1 and "start(5.000000) > max(3.000000)" in r.stderr


def test_product_host_sources_never_touch_the_oracle():
    """the product's host layer knows nothing about the oracle (the shim lives in tests/)"""
    for f in os.listdir(HOST):
        if f.endswith((".c", ".h")) or f == "Makefile":
            text = open(os.path.join(HOST, f)).read()
            assert "orc_" not in text and "oracle" not in text.lower(), f
