#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference
built as oracle/_ref (run `make -C oracle ref` first; needs /root/reference).

What is produced (all small JSON, floats as repr strings so they round-trip):
  eval_<model>.json   apps/eval_main.c output (prob, prior; %.15e) of the reference's
                      calc_model on a seeded grid of parameter vectors, with the
                      params rows and the data table used.
  eval_simplesin5.json  same for apps/simplesin5.c, which does not compile as shipped
                      (SURVEY.md D1): a copy under /tmp is sed-patched (drop the write to
                      the non-existent m->model, loop bound -> m->data->size1) and built
                      with the same recipe.  No reference source enters the repo.
  c1_phases.json      config C1 in small: simplesin on tests/testlc.dat, N_BETA=4,
                      BURN_IN_ITERATIONS=1000, MAX_ITERATIONS=3000, GSL_RNG_SEED=7, one
                      thread: calibration_results after each of calibrate_first /
                      calibrate_rest, sha256 + head/tail of every dump of `run`, and the
                      evidence `analyse` prints.
  c1_{circular,logistic,uniform}_phases.json   same with -DCIRCULAR_PARAMS=3,
                      -DPROPOSAL_LOGISTIC, -DPROPOSAL_UNIFORM.
  c4_phases.json      pulse_vrot (a model with a prior) on a synthetic spectrum, N_BETA=3.
  c2_phases.json      normal (data-free), N_BETA=5.
  testlc.dat          the reference's data fixture for config C1 (tests/testlc.dat).

Usage: python tests/golden/make_golden.py
"""
import hashlib
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
from oracle_binding import ref_exe, write_params_file, write_data_file, ref_eval  # noqa: E402

REF = "/root/reference"


def frepr(a):
    return [repr(float(x)) for x in np.asarray(a).ravel()]


def lightcurve(n, rng, span=50.0):
    x = np.sort(rng.uniform(0, span, n))
    y = 1.3 * np.sin(2 * np.pi * (7.25 * x + 0.31)) + 0.2 + rng.normal(0, 0.5, n)
    return np.stack([x, y], axis=1)


def pulse_spectrum(n, rng, vrot=True):
    nu = np.linspace(90, 110, n)
    tau = 0.5

    def lor(d, h):
        return h / (1 + (2 * np.pi * d * tau) ** 2)
    y = lor(98 - nu, 5.0) + lor(103 - nu, 3.0)
    if vrot:
        y = y + lor(103 - nu - 0.4, 3.0) + lor(103 - nu + 0.4, 3.0)
    y = y + 0.05
    return np.stack([nu, y * rng.exponential(1.0, n)], axis=1)


MODELS = {
    "simplesin": dict(
        rows=[(1.0, 0.0, 3.0, "amplitude", -1.0), (7.25, 4.0, 10.0, "frequency", 0.001),
              (0.31, 0.0, 1.0, "phase", -1.0), (0.2, -1.0, 1.0, "offset", -1.0)],
        data=lambda rng: lightcurve(257, rng)),
    "simplesin2": dict(
        rows=[(1.0, 0.0, 3.0, "amplitude", -1.0), (7.25, 4.0, 10.0, "frequency", 0.001)],
        data=lambda rng: lightcurve(130, rng)),
    "normal": dict(
        rows=[(100.0, -10.0, 10000.0, "x", -1.0)],
        data=lambda rng: np.array([[0.0, 0.0], [1.0, 1.0]])),
    "pulse_vrot": dict(
        rows=[(0.5, 0.01, 5.0, "lifetime", -1.0), (0.0, -10.0, 10.0, "p1", -1.0),
              (0.4, 0.0, 2.0, "vrot", -1.0), (98.0, 95.0, 100.0, "f1", -1.0),
              (5.0, 0.01, 20.0, "h1", -1.0), (103.0, 100.0, 106.0, "f2", -1.0),
              (3.0, 0.01, 20.0, "h2", -1.0)],
        data=lambda rng: pulse_spectrum(300, rng)),
    "pulse": dict(
        rows=[(0.5, 0.01, 5.0, "lifetime", -1.0), (0.0, -10.0, 10.0, "p1", -1.0),
              (98.0, 95.0, 100.0, "f1", -1.0), (5.0, 0.01, 20.0, "h1", -1.0),
              (103.0, 100.0, 106.0, "f2", -1.0), (3.0, 0.01, 20.0, "h2", -1.0)],
        data=lambda rng: pulse_spectrum(200, rng, vrot=False)),
    "bernoulli_example": dict(
        rows=[(0.1, -5.0, 5.0, "b0", -1.0), (0.5, -5.0, 5.0, "b1", -1.0), (-0.3, -5.0, 5.0, "b2", -1.0)],
        data=lambda rng: (lambda X: np.column_stack([
            (rng.uniform(size=150) < 1 / (1 + np.exp(-(0.1 + X @ np.array([0.5, -0.3]))))).astype(float), X]))(
                rng.normal(size=(150, 2)))),
}


def param_grid(rows, rng, n=24):
    lo = np.array([r[1] for r in rows])
    hi = np.array([r[2] for r in rows])
    g = rng.uniform(lo, hi, size=(n, len(rows)))
    g[0] = [r[0] for r in rows]
    return g


def eval_fixture(model, exe_model, spec, seed):
    rng = np.random.default_rng(seed)
    data = spec["data"](rng)
    grid = param_grid(spec["rows"], rng)
    if model == "normal":
        grid[1:7, 0] = [1.0, np.e, 7.3, 20.0, -3.0, 9000.0]
    with tempfile.TemporaryDirectory() as d:
        write_params_file(os.path.join(d, "params"), spec["rows"])
        write_data_file(os.path.join(d, "data"), data)
        out = ref_eval(exe_model, d, grid)
    assert out.shape == (len(grid), 2), out.shape
    return dict(model=model, rows=[list(r) for r in spec["rows"]], n_cols=int(data.shape[1]),
                data=frepr(data), params=frepr(grid), n_vectors=len(grid),
                prob=frepr(out[:, 0]), prior=frepr(out[:, 1]),
                source=f"oracle/_ref/eval_{exe_model}.exe (reference apps/eval_main.c + apps/{model}.c)")


def build_patched_simplesin5():
    """apps/simplesin5.c with the two stale lines fixed, compiled from a /tmp copy."""
    tmp = tempfile.mkdtemp(prefix="ss5_")
    src = open(os.path.join(REF, "apps", "simplesin5.c")).read()
    src = src.replace("\tgsl_vector_set(m->model, i, y);\n", "")
    src = src.replace("m->x_dat->size", "m->data->size1")
    os.makedirs(os.path.join(tmp, "apps"))
    open(os.path.join(tmp, "apps", "simplesin5fix.c"), "w").write(src)
    out = os.path.join(ROOT, "oracle", "_ref", "eval_simplesin5fix.exe")
    engine = [os.path.join(REF, "src", f) for f in sorted(os.listdir(os.path.join(REF, "src"))) if f.endswith(".c")]
    subprocess.run(["gcc", "-I", os.path.join(REF, "src"), "-I", os.path.join(ROOT, "apemost_b200", "compat"),
                    "-O3", "-std=c99", "-fopenmp", "-fPIC", "-ansi", "-pedantic", "-DWITHOUT_GARBAGE_COLLECTOR",
                    os.path.join(tmp, "apps", "simplesin5fix.c"), os.path.join(REF, "apps", "eval_main.c"), *engine,
                    os.path.join(ROOT, "oracle", "_build", "gsl_compat.o"), "-lm", "-lgomp", "-o", out],
                   check=True, stderr=subprocess.DEVNULL)
    shutil.rmtree(tmp)
    return out


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def phases_fixture(model, rows, data_file, data, cfg, ccflags_extra="", suffix="_pin1", engine_opts=None):
    """Run calibrate_first / calibrate_rest / run / analyse of the reference and record the
    files they leave behind."""
    flags = " ".join(f"-D{k}={v}" for k, v in cfg.items() if k != "GSL_RNG_SEED")
    exe = ref_exe(model, ccflags=(flags + " " + ccflags_extra).strip(), suffix=suffix)
    out = dict(config=cfg, ccflags_extra=ccflags_extra, engine_opts=engine_opts or {},
               rows=[list(r) for r in rows], model=model, data_file=data_file,
               data=None if data_file else frepr(data), n_cols=2, phases={}, dumps={})
    with tempfile.TemporaryDirectory() as d:
        write_params_file(os.path.join(d, "params"), rows)
        if data_file:
            shutil.copy(os.path.join(HERE, data_file), os.path.join(d, "data"))
        else:
            write_data_file(os.path.join(d, "data"), data)
        env = dict(os.environ, GSL_RNG_SEED=str(cfg["GSL_RNG_SEED"]), OMP_NUM_THREADS="1")
        for phase in ("calibrate_first", "calibrate_rest", "run"):
            subprocess.run([exe, phase], cwd=d, env=env, check=True, capture_output=True)
            out["phases"][phase] = open(os.path.join(d, "calibration_results")).read()
        for f in sorted(os.listdir(d)):
            if f.endswith(".dump") and "acceptance" not in f:
                lines = open(os.path.join(d, f)).read().splitlines()
                out["dumps"][f] = dict(sha256=sha(os.path.join(d, f)), n_lines=len(lines),
                                       head=lines[:5], tail=lines[-5:])
        r = subprocess.run([exe, "analyse"], cwd=d, env=env, check=True, capture_output=True, text=True)
        m = re.search(r"Model probability ln\(p\(D\|M, I\)\): \[about 10\^(-?\d+)\] (-?[\d.]+)", r.stdout)
        out["evidence_line"] = m.group(0)
        out["evidence"] = m.group(2)
        # every other file the four phases leave behind (params_suggested, calibration_summary,
        # calibration_progress.data, acceptance_rate.dump[.gnuplot], <name>.histogram,
        # marginal_distributions.gnuplot) and what analyse prints: pins the host layer's formats
        out["files"] = {}
        for f in sorted(os.listdir(d)):
            if f in ("params", "data", "gmon.out") or f in out["dumps"]:
                continue
            text = open(os.path.join(d, f)).read()
            lines = text.splitlines()
            out["files"][f] = dict(sha256=sha(os.path.join(d, f)), n_lines=len(lines),
                                   text=text if len(text) <= 6000 else None,
                                   head=lines[:3], tail=lines[-3:])
        out["analyse_stdout"] = r.stdout
    return out


C1_ROWS = [(1.0, 0.0, 3.0, "amplitude", -1.0), (15.2, 4.0, 24.0, "frequency", 0.001),
           (0.25, 0.0, 1.0, "phase", -1.0), (0.0, -1.0, 1.0, "offset", -1.0)]


def all_phase_fixtures(wanted=None):
    """name -> fixture for every phase fixture (or the wanted ones): each entry below is a recipe"""
    fx = {}
    small = dict(N_BETA=4, BURN_IN_ITERATIONS=1000, MAX_ITERATIONS=3000, GSL_RNG_SEED=7)
    # config C1 in small: simplesin on the reference's own light curve
    fx["c1_phases"] = lambda: phases_fixture("simplesin", C1_ROWS, "testlc.dat", None, small)
    # circular phase parameter (CIRCULAR_PARAMS lists parameter 3 = phase)
    fx["c1_circular_phases"] = lambda: phases_fixture(
        "simplesin", C1_ROWS, "testlc.dat", None, dict(small, GSL_RNG_SEED=11),
        ccflags_extra="-DCIRCULAR_PARAMS=3", suffix="_pin_circ", engine_opts=dict(circular_mask=4))
    # alternative proposal distributions
    fx["c1_logistic_phases"] = lambda: phases_fixture(
        "simplesin", C1_ROWS, "testlc.dat", None, dict(small, GSL_RNG_SEED=3),
        ccflags_extra="-DPROPOSAL_LOGISTIC", suffix="_pin_logi", engine_opts=dict(proposal=1))
    fx["c1_uniform_phases"] = lambda: phases_fixture(
        "simplesin", C1_ROWS, "testlc.dat", None, dict(small, GSL_RNG_SEED=5),
        ccflags_extra="-DPROPOSAL_UNIFORM", suffix="_pin_unif", engine_opts=dict(proposal=2))
    # config C4 in small: pulse_vrot (has a prior: pins the stale-prior and swap-ratio quirks)
    rng = np.random.default_rng(4242)
    fx["c4_phases"] = lambda: phases_fixture(
        "pulse_vrot", MODELS["pulse_vrot"]["rows"], None, pulse_spectrum(200, rng),
        dict(N_BETA=3, BURN_IN_ITERATIONS=1000, MAX_ITERATIONS=2000, GSL_RNG_SEED=9), suffix="_pin4")
    # config C2 in small: normal (data-free)
    fx["c2_phases"] = lambda: phases_fixture(
        "normal", [(20.0, 0.0, 60.0, "x", -1.0)], None, np.array([[0.0, 0.0], [1.0, 1.0]]),
        dict(N_BETA=3, BURN_IN_ITERATIONS=1000, MAX_ITERATIONS=4000, BETA_0=0.5, GSL_RNG_SEED=2),
        suffix="_pin2b", engine_opts=dict(beta_0=0.5))
    # -DADAPT: long enough for the 1 % rescalings (counter sums >= 20000) and a counter reset
    # (> 100000) to happen; -DRANDOMSWAP: one more uniform per round
    fx["c1_adapt_phases"] = lambda: phases_fixture(
        "simplesin", C1_ROWS, "testlc.dat", None, dict(small, MAX_ITERATIONS=30000, GSL_RNG_SEED=13),
        ccflags_extra="-DADAPT", suffix="_pin_adapt", engine_opts=dict(adapt=0.5))
    fx["c1_randomswap_phases"] = lambda: phases_fixture(
        "simplesin", C1_ROWS, "testlc.dat", None, dict(small, GSL_RNG_SEED=17),
        ccflags_extra="-DRANDOMSWAP", suffix="_pin_rswap", engine_opts=dict(random_swap=1))
    # -DCALIBRATE_ALTERNATE: assess_acceptance_rate + markov_chain_calibrate_alt replace _orig
    # (pins the host layer's apm_calibrate_alt.c through tests/test_host_cpu.py).  The reference's
    # alternate calibrator often gives up on the hot chains ("iteration limit reached", exit 1), so
    # the fixture calibrates chains 0 and 1 with it and only burns the others in
    # (-DSKIP_CALIBRATE_ALLCHAINS), with a seed for which it converges.
    fx["c1_altcal_phases"] = lambda: phases_fixture(
        "simplesin", C1_ROWS, "testlc.dat", None, dict(small, GSL_RNG_SEED=23),
        ccflags_extra="-DCALIBRATE_ALTERNATE -DSKIP_CALIBRATE_ALLCHAINS", suffix="_pin_altskip",
        engine_opts=dict(host_only=1))
    # -DCALIBRATE_MULTILIN: markov_chain_calibrate_multilinear_regression for every chain (pins
    # apm_calibrate_multilin.c and the host-side uniform draws).  On the hottest chain the
    # reference's regression walks one step width below zero -- kept: it is what the reference does.
    fx["c1_multilin_phases"] = lambda: phases_fixture(
        "simplesin", C1_ROWS, "testlc.dat", None, dict(small, GSL_RNG_SEED=29),
        ccflags_extra="-DCALIBRATE_MULTILIN", suffix="_pin_multilin", engine_opts=dict(host_only=1))
    # -DCALIBRATE_QUADRATIC: markov_chain_calibrate_quadratic + markov_chain_calibrate_linear_regression
    # (pins apm_calibrate_quadratic.c).  The reference's parabola search leaves wild step widths
    # (1e-14 .. 1e21 of the range) and for most seeds ends in an assertion or a GSL range error on a
    # hot chain, so the fixture calibrates chains 0 and 1 with it (-DSKIP_CALIBRATE_ALLCHAINS) with a
    # seed for which both survive.
    fx["c1_quadratic_phases"] = lambda: phases_fixture(
        "simplesin", C1_ROWS, "testlc.dat", None, dict(small, GSL_RNG_SEED=4),
        ccflags_extra="-DCALIBRATE_QUADRATIC -DSKIP_CALIBRATE_ALLCHAINS", suffix="_pin_quad",
        engine_opts=dict(host_only=1))
    return {name: recipe() for name, recipe in fx.items() if wanted is None or name in wanted}


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    shutil.copy(os.path.join(REF, "tests", "testlc.dat"), os.path.join(HERE, "testlc.dat"))
    for i, (model, spec) in enumerate(MODELS.items()):
        fx = eval_fixture(model, model, spec, 1000 + i)
        json.dump(fx, open(os.path.join(HERE, f"eval_{model}.json"), "w"), indent=0)
        print("wrote eval_%s.json" % model)
    build_patched_simplesin5()
    spec5 = dict(rows=[(1.3, 0.0, 3.0, "amplitude", -1.0), (7.25, 4.0, 10.0, "frequency", 0.001),
                       (1.9, 0.0, 6.283185307179586, "phase", -1.0), (0.2, -1.0, 1.0, "offset", -1.0)],
                 data=lambda rng: lightcurve(257, rng, span=1000.0))
    fx = eval_fixture("simplesin5", "simplesin5fix", spec5, 1100)
    fx["source"] = ("apps/simplesin5.c patched in /tmp (drop m->model write, loop bound m->data->size1; "
                    "SURVEY.md D1) + apps/eval_main.c, built with the oracle/_ref recipe")
    json.dump(fx, open(os.path.join(HERE, "eval_simplesin5.json"), "w"), indent=0)
    print("wrote eval_simplesin5.json")
    for name, fx in all_phase_fixtures().items():
        json.dump(fx, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
        print("wrote %s.json" % name)


def only(names):
    """regenerate some of the phase fixtures: python make_golden.py c1_quadratic_phases ..."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "_build", "gsl_compat.o")],
                   check=True)
    fxs = all_phase_fixtures(set(names))
    for name in names:
        json.dump(fxs[name], open(os.path.join(HERE, name + ".json"), "w"), indent=1)
        print("wrote %s.json" % name)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        only(sys.argv[1:])
    else:
        main()
