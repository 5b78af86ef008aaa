"""The product's C host layer over the real engine (GPU box): `make <model>.exe`, then the
reference's phases in a scratch working directory.  The same flow is driven in Python over the CPU
oracle in PHILOX mode (tests/pt_flow.py, itself byte-pinned to the reference in MT19937 mode), so
the files must agree number for number: same accept / swap / rescale decisions, values to 1e-9
(the device's log / sqrt / cos in the Box-Muller transform differ from libm by an ulp, so the
text is not bit-identical); prob-chain dumps carry 7 digits of a sum whose order differs."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import pt_flow
from oracle_binding import Oracle, RNG_PHILOX, write_data_file, write_params_file

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
HOST = os.path.join(ROOT, "apemost_b200", "host")


def make(target, ccflags, out):
    os.makedirs(out, exist_ok=True)
    subprocess.run(["make", "-s", "-C", HOST, f"OUT={out}", f"CCFLAGS={ccflags}", os.path.join(out, target)],
                   check=True)
    return os.path.join(out, target)


def setup_workdir(wd, fx):
    os.makedirs(wd, exist_ok=True)
    write_params_file(os.path.join(wd, "params"), [tuple(r) for r in fx["rows"]])
    if fx["data_file"]:
        data = np.loadtxt(os.path.join(GOLDEN, fx["data_file"]))
        open(os.path.join(wd, "data"), "wb").write(open(os.path.join(GOLDEN, fx["data_file"]), "rb").read())
    else:
        data = np.array(fx["data"], dtype=float).reshape(-1, fx["n_cols"])
        write_data_file(os.path.join(wd, "data"), data)
    return data


def numbers(path):
    return np.array(open(path).read().split(), dtype=float)


@pytest.mark.parametrize("name", ["c1_phases", "c4_phases", "c2_phases"])
def test_phases_on_gpu_equal_oracle_flow(name, tmp_path):
    fx = json.load(open(os.path.join(GOLDEN, name + ".json")))
    cfg, opts = fx["config"], fx["engine_opts"]
    seed = cfg["GSL_RNG_SEED"]
    flags = " ".join(f"-D{k}={v}" for k, v in cfg.items() if k != "GSL_RNG_SEED")
    exe = make(fx["model"] + ".exe", flags, str(tmp_path / "bin"))
    wd_gpu, wd_cpu = str(tmp_path / "gpu"), str(tmp_path / "cpu")
    data = setup_workdir(wd_gpu, fx)
    setup_workdir(wd_cpu, fx)
    rows = [tuple(r) for r in fx["rows"]]
    env = dict(os.environ, GSL_RNG_SEED=str(seed))

    def oracle():
        e = Oracle(fx["model"], 1, cfg["N_BETA"], n_par=len(rows), seed=seed, rng=RNG_PHILOX)
        e.set_data(data)
        return e

    burn = cfg["BURN_IN_ITERATIONS"]
    steps = {
        "calibrate_first": lambda: pt_flow.calibrate_first(oracle(), rows, wd_cpu, burn_in_iterations=burn),
        "calibrate_rest": lambda: pt_flow.calibrate_rest(oracle(), rows, wd_cpu, beta_0=opts.get("beta_0", -0.001),
                                                         burn_in_iterations=burn),
        "run": lambda: pt_flow.run(oracle(), rows, wd_cpu, cfg["MAX_ITERATIONS"]),
    }
    for phase, cpu_phase in steps.items():
        subprocess.run([exe, phase], cwd=wd_gpu, env=env, check=True, capture_output=True)
        cpu_phase()
        got = numbers(os.path.join(wd_gpu, "calibration_results"))
        want = numbers(os.path.join(wd_cpu, "calibration_results"))
        assert got.shape == want.shape, phase
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-300, err_msg=phase)
    for f in sorted(os.listdir(wd_cpu)):
        if f.endswith(".prob.dump"):
            a, b = numbers(os.path.join(wd_gpu, f)), numbers(os.path.join(wd_cpu, f))
            assert a.shape == b.shape, f
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-300, err_msg=f)
        elif f.startswith("prob-chain"):
            a, b = numbers(os.path.join(wd_gpu, f)), numbers(os.path.join(wd_cpu, f))
            assert a.shape == b.shape, f
            np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-12, err_msg=f)
    r = subprocess.run([exe, "analyse"], cwd=wd_gpu, env=env, check=True, capture_output=True, text=True)
    m = re.search(r"Model probability ln\(p\(D\|M, I\)\): \[about 10\^(-?\d+)\] (-?[\d.]+)", r.stdout)
    beta = pt_flow.read_calibration_results(os.path.join(wd_cpu, "calibration_results"), cfg["N_BETA"], len(rows))[0]
    mean_dl = [np.mean(numbers(os.path.join(wd_cpu, f"prob-chain{k}.dump")).reshape(-1, 2)[:, 1])
               for k in range(cfg["N_BETA"])]
    assert abs(float(m.group(2)) - pt_flow.evidence(beta, mean_dl)) < 1e-4
    for f in ("params_suggested", "calibration_summary", "acceptance_rate.dump.gnuplot", "run_statistics",
              "marginal_distributions.gnuplot"):
        assert os.path.getsize(os.path.join(wd_gpu, f)) > 0, f


def test_eval_exe_matches_reference_fixture(tmp_path):
    """eval_<model>.exe (reference apps/eval_main.c) answered by the GPU: 1e-12 against what the
    unmodified reference printed for the same parameter vectors"""
    for model in ("simplesin", "pulse_vrot", "bernoulli_example"):
        fx = json.load(open(os.path.join(GOLDEN, f"eval_{model}.json")))
        exe = make(f"eval_{model}.exe", "-DN_BETA=1", str(tmp_path / "bin"))
        wd = str(tmp_path / model)
        os.makedirs(wd)
        write_params_file(os.path.join(wd, "params"), [tuple(r) for r in fx["rows"]])
        data = np.array(fx["data"], dtype=float).reshape(-1, fx["n_cols"])
        write_data_file(os.path.join(wd, "data"), data)
        vectors = np.array(fx["params"], dtype=float).reshape(fx["n_vectors"], -1)
        text = "\n".join(" ".join(repr(float(v)) for v in row) for row in vectors) + "\n"
        r = subprocess.run([exe], cwd=wd, input=text, capture_output=True, text=True, check=True)
        out = np.array([l.split() for l in r.stdout.splitlines() if re.match(r"^-?\d", l)], dtype=float)
        np.testing.assert_allclose(out[:, 0], np.array(fx["prob"], dtype=float), rtol=1e-12)
        np.testing.assert_allclose(out[:, 1], np.array(fx["prior"], dtype=float), rtol=1e-12, atol=1e-300)


def test_benchmark_exe_reproduces_and_matches_fixture(tmp_path):
    """benchmark_<model>.exe (reference apps/benchmark_main.c): n + p evaluations at the params
    file's start values, every one bit-identical to the first; the printed prob against the
    reference's eval fixture (its first vector is the start vector)"""
    fx = json.load(open(os.path.join(GOLDEN, "eval_simplesin.json")))
    exe = make("benchmark_simplesin.exe", "-DN_BETA=1", str(tmp_path / "bin"))
    wd = str(tmp_path / "wd")
    os.makedirs(wd)
    write_params_file(os.path.join(wd, "params"), [tuple(r) for r in fx["rows"]])
    write_data_file(os.path.join(wd, "data"), np.array(fx["data"], dtype=float).reshape(-1, fx["n_cols"]))
    r = subprocess.run([exe, "1000", "9000"], cwd=wd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    m = re.search(r"^10000 model evaluations in .* prob = (\S+)$", r.stdout, flags=re.M)
    assert m, r.stdout
    assert abs(float(m.group(1)) / float(fx["prob"][0]) - 1) < 1e-12
    r = subprocess.run([exe], cwd=wd, capture_output=True, text=True)
    assert r.returncode == 1 and "SYNOPSIS" in r.stderr


@pytest.mark.parametrize("fixture", ["c1_altcal_phases", "c1_multilin_phases", "c1_quadratic_phases"])
def test_alternate_calibrator_on_gpu_equals_oracle_flow(tmp_path, fixture):
    """-DCALIBRATE_ALTERNATE (assess_acceptance_rate + markov_chain_calibrate_alt, reference
    src/markov_chain.c:117-224, src/markov_chain_calibrate.c:927-1037) and -DCALIBRATE_MULTILIN
    (markov_chain_calibrate_multilinear_regression, :33-237, with its host-side uniform draws through
    apm_gpu_host_uniform) in the host layer: the product executable on the GPU against the very same
    host sources linked to the CPU oracle in PHILOX mode (whose MT19937 twin is byte-pinned to a
    reference build with the same flag by test_host_cpu.py)"""
    from test_host_cpu import build_host_over_oracle
    fx = json.load(open(os.path.join(GOLDEN, fixture + ".json")))
    cfg = fx["config"]
    flags = [f"-D{k}={v}" for k, v in cfg.items() if k != "GSL_RNG_SEED"] + fx["ccflags_extra"].split()
    exe_gpu = make("simplesin.exe", " ".join(flags), str(tmp_path / "bin"))
    exe_cpu = build_host_over_oracle(fixture + "_philox", "simplesin", flags)
    out = {}
    for name, exe, extra in (("gpu", exe_gpu, {}), ("cpu", exe_cpu, {"APM_TEST_ORACLE_RNG": "philox"})):
        wd = str(tmp_path / name)
        setup_workdir(wd, fx)
        env = dict(os.environ, GSL_RNG_SEED="5", **extra)
        codes, said = [], []
        for phase in ("calibrate_first", "calibrate_rest"):
            r = subprocess.run([exe, phase], cwd=wd, env=env, capture_output=True, text=True)
            codes.append((phase, r.returncode, r.stderr.strip().splitlines()[-1:] if r.returncode else []))
            said += re.findall(r"^ +\d+ iterations$", r.stdout, flags=re.M)  # the regression calibrator's log
            if r.returncode != 0:
                # the reference's alternate calibrator gives up easily ("iteration limit reached",
                # exit 1, src/markov_chain_calibrate.c:1017-1020); then both must give up alike
                assert r.returncode == 1 and "iteration limit reached" in r.stderr, r.stderr[-500:]
                break
        progress = os.path.join(wd, "calibration_progress.data")  # written by calibrate_alt only
        out[name] = (codes, numbers(os.path.join(wd, "calibration_results")),
                     open(progress).read() if os.path.exists(progress) else "\n".join(said))
    assert out["gpu"][0] == out["cpu"][0] and out["gpu"][0][0][1] == 0, out["gpu"][0]
    np.testing.assert_allclose(out["gpu"][1], out["cpu"][1], rtol=1e-9, atol=1e-300)
    assert out["gpu"][2] == out["cpu"][2] and len(out["gpu"][2]) > 0


@pytest.mark.parametrize("name", ["c1_phases", "c4_phases"])
def test_analyse_from_device_accumulators_equals_analyse_from_dumps(name, tmp_path):
    """SURVEY.md 8 f1 on the device: the 200-bin histograms and the batch means `run` accumulates in
    its kernels (apm_gpu_set_marginals) give the same <name>.histogram files, byte for byte, and the
    same error estimates as analyse's pass over the parameter dumps of the same run"""
    from test_host_cpu import check_analyse_from_accumulators
    fx = json.load(open(os.path.join(GOLDEN, name + ".json")))
    cfg = fx["config"]
    flags = " ".join(f"-D{k}={v}" for k, v in cfg.items() if k != "GSL_RNG_SEED")
    exe = make(fx["model"] + ".exe", flags, str(tmp_path / "bin"))
    wd = str(tmp_path / "wd")
    setup_workdir(wd, fx)
    env = dict(os.environ, GSL_RNG_SEED=str(cfg["GSL_RNG_SEED"]))
    for phase in ("calibrate_first", "calibrate_rest"):
        subprocess.run([exe, phase], cwd=wd, env=env, check=True, capture_output=True)
    check_analyse_from_accumulators(exe, wd, env, [r[3] for r in fx["rows"]])


def test_several_ensembles_side_by_side(tmp_path):
    """N_ENSEMBLES=3: one working directory per ensemble; ensemble 0 (chain ids 0..n_beta-1,
    ensemble id 0) repeats the single-ensemble run exactly, the others differ"""
    fx = json.load(open(os.path.join(GOLDEN, "c1_phases.json")))
    base = "-DN_BETA=4 -DBURN_IN_ITERATIONS=600 -DMAX_ITERATIONS=1500"
    exe1 = make("simplesin.exe", base, str(tmp_path / "bin1"))
    exe3 = make("simplesin.exe", base + " -DN_ENSEMBLES=3", str(tmp_path / "bin3"))
    wd1, wd3 = str(tmp_path / "one"), str(tmp_path / "three")
    setup_workdir(wd1, fx)
    setup_workdir(wd3, fx)
    env = dict(os.environ, GSL_RNG_SEED="21")
    for phase in ("calibrate_first", "calibrate_rest", "run"):
        subprocess.run([exe1, phase], cwd=wd1, env=env, check=True, capture_output=True)
        subprocess.run([exe3, phase], cwd=wd3, env=env, check=True, capture_output=True)
    for f in ("calibration_results", "amplitude-chain-0.prob.dump", "prob-chain3.dump", "acceptance_rate.dump"):
        assert open(os.path.join(wd1, f)).read() == open(os.path.join(wd3, "ens0", f)).read(), f
    a = open(os.path.join(wd3, "ens1", "frequency-chain-0.prob.dump")).read()
    b = open(os.path.join(wd3, "ens2", "frequency-chain-0.prob.dump")).read()
    assert a != b and a != open(os.path.join(wd3, "ens0", "frequency-chain-0.prob.dump")).read()
    assert len(a.splitlines()) == 1500
    r = subprocess.run([exe3, "analyse"], cwd=wd3, env=env, check=True, capture_output=True, text=True)
    assert "3 independent ensembles" in r.stdout


def test_user_model_device_against_host_plugin(tmp_path):
    """apps/multisin.{c,cuh}: a model file in APEMoST's plugin style with its __device__ counterpart,
    built as APM_MODEL_USER.  K = 1 is the reference's simplesin, so the device AND the host half
    must reproduce what the unmodified reference printed (1e-12); for K = 3 device and host must
    agree with each other; `check` reports the same; and the four phases run."""
    fx = json.load(open(os.path.join(GOLDEN, "eval_simplesin.json")))
    exe = make("eval_multisin.exe", "-DN_BETA=1", str(tmp_path / "bin"))
    wd = str(tmp_path / "k1")
    os.makedirs(wd)
    write_params_file(os.path.join(wd, "params"), [tuple(r) for r in fx["rows"]])
    data = np.array(fx["data"], dtype=float).reshape(-1, fx["n_cols"])
    write_data_file(os.path.join(wd, "data"), data)
    vectors = np.array(fx["params"], dtype=float).reshape(fx["n_vectors"], -1)
    text = "\n".join(" ".join(repr(float(v)) for v in row) for row in vectors) + "\n"

    def evaluate(args, cwd, stdin):
        r = subprocess.run([exe, *args], cwd=cwd, input=stdin, capture_output=True, text=True, check=True)
        return np.array([l.split() for l in r.stdout.splitlines() if re.match(r"^-?\d", l)], dtype=float)[:, 0]

    want = np.array(fx["prob"], dtype=float)
    np.testing.assert_allclose(evaluate([], wd, text), want, rtol=1e-12)
    np.testing.assert_allclose(evaluate(["--host"], wd, text), want, rtol=1e-12)

    # K = 3 on a 5000-row light curve
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0, 300, 5000))
    truth = [1.0, 3.1, 0.2, 0.5, 7.7, 0.6, 0.3, 11.3, 0.9, 0.1]
    y = sum(truth[3 * k] * np.sin(2 * np.pi * (truth[3 * k + 1] * x + truth[3 * k + 2])) for k in range(3)) + truth[9]
    y = y + rng.normal(0, 0.5, x.size)
    wd3 = str(tmp_path / "k3")
    os.makedirs(wd3)
    rows = []
    for k in range(3):
        rows += [(truth[3 * k], 0.0, 3.0, f"amplitude{k}", -1.0), (truth[3 * k + 1], 1.0, 15.0, f"frequency{k}", 1e-4),
                 (truth[3 * k + 2], 0.0, 1.0, f"phase{k}", -1.0)]
    rows.append((truth[9], -1.0, 1.0, "offset", -1.0))
    write_params_file(os.path.join(wd3, "params"), rows)
    write_data_file(os.path.join(wd3, "data"), np.stack([x, y], axis=1))
    vec = np.array(truth)[None, :] + rng.normal(0, 1e-3, (16, 10))
    text3 = "\n".join(" ".join(repr(float(v)) for v in row) for row in vec) + "\n"
    np.testing.assert_allclose(evaluate([], wd3, text3), evaluate(["--host"], wd3, text3), rtol=1e-12)

    main = make("multisin.exe", "-DN_BETA=4 -DBURN_IN_ITERATIONS=2000 -DMAX_ITERATIONS=2000 -DCIRCULAR_PARAMS=3,6,9",
                str(tmp_path / "bin"))
    r = subprocess.run([main, "check"], cwd=wd3, capture_output=True, text=True, check=True)
    m = re.search(r"relative difference: ([\d.e+-]+)", r.stdout)
    assert m and float(m.group(1)) < 1e-12, r.stdout[-400:]
    for phase in ("calibrate_first", "calibrate_rest", "run", "analyse"):
        r = subprocess.run([main, phase], cwd=wd3, capture_output=True, text=True, check=True)
    assert "Model probability ln(p(D|M, I))" in r.stdout
    assert len(open(os.path.join(wd3, "frequency1-chain-0.prob.dump")).read().splitlines()) == 2000
    # the posterior mean of the first frequency sits at the truth
    f0 = np.loadtxt(os.path.join(wd3, "frequency0-chain-0.prob.dump"))
    assert abs(f0[500:].mean() - truth[1]) < 5e-3
