"""Pins the CPU oracle (oracle/apm_oracle.c) to the reference.

Fixtures in tests/golden/ were produced by the UNMODIFIED reference built as
oracle/_ref (tests/golden/make_golden.py).  Floating point: calc_model values are
compared at 4e-15 relative (the reference prints %.15e = 16 significant digits);
complete phases in MT19937 mode must reproduce the reference's files byte for byte.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import pt_flow
from oracle_binding import Oracle, RNG_MT19937, RNG_PHILOX, oracle_lib, philox, evidence

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EVAL_MODELS = ["simplesin", "simplesin5", "simplesin2", "normal", "pulse_vrot", "pulse", "bernoulli_example"]
ENGINE_NAME = {"bernoulli_example": "bernoulli"}


def load(name):
    return json.load(open(os.path.join(GOLDEN, name + ".json")))


def fx_arrays(fx):
    n_par = len(fx["rows"])
    data = np.array(fx["data"], dtype=float).reshape(-1, fx["n_cols"])
    params = np.array(fx["params"], dtype=float).reshape(-1, n_par)
    return data, params, np.array(fx["prob"], dtype=float), np.array(fx["prior"], dtype=float)


def test_manual_known_answer():
    """reference doc/manual.rst:189-213"""
    eng = Oracle("simplesin", 1, 1)
    eng.set_data(np.array([[101, 0.67], [102, 1.01], [103, 0.79], [104, 1.34]]))
    prob, prior = eng.eval([[1, 0.2, 1, 0]])
    assert "%.15e" % prob[0] == "-1.480898044165363e+01"
    assert "%.15e" % prior[0] == "0.000000000000000e+00"


@pytest.mark.parametrize("model", EVAL_MODELS)
def test_calc_model_matches_reference_eval(model):
    fx = load("eval_" + model)
    data, params, prob_ref, prior_ref = fx_arrays(fx)
    eng = Oracle(ENGINE_NAME.get(model, model), 1, 1, n_par=params.shape[1])
    eng.set_data(data)
    prob, prior = eng.eval(params)
    np.testing.assert_allclose(prob, prob_ref, rtol=4e-15, atol=0)
    np.testing.assert_allclose(prior, prior_ref, rtol=4e-15, atol=0)


def test_normal_known_values():
    """SURVEY.md 8c probe values for apps/normal.c"""
    eng = Oracle("normal", 1, 1)
    eng.set_data(np.zeros((2, 2)))
    prob, _ = eng.eval([[1.0], [np.e], [7.3], [20.0]])
    np.testing.assert_allclose(prob, [0.0, 10.0, 9.998017252810813, 9.714876922707774], rtol=1e-15)


def test_mod_double():
    """reference tests/tests.c:152-160 (1e-3 relative, tests.c:39-49)"""
    lib = oracle_lib()
    cases = [(3.14, 3.00, 0.14), (3.14, 1.30, 0.54), (-3.14, 1.30, 0.76), (0, 1.30, 0.00),
             (6000.3214, 1.1324, 0.8662), (-6000.3214, 1.1324, 0.2662)]
    for x, d, want in cases:
        got = lib.orc_mod_double(x, d)
        assert abs(got - want) <= 1e-3 * max(abs(want), 1e-300) + 1e-12, (x, d, got, want)


def test_chebyshev_ladder_manual():
    """reference doc/manual.rst:419-440: N_BETA=20, automatic beta_0 = 0.013294"""
    lib = oracle_lib()
    want = {0: 1.0, 1: 0.993271, 2: 0.973269, 3: 0.940538, 4: 0.895972, 9: 0.547388,
            10: 0.465906, 17: 0.040026, 18: 0.020023, 19: 0.013294}
    for i, b in want.items():
        assert abs(lib.orc_get_chain_beta(i, 20, 0.013294) - b) < 1.1e-6  # beta_0 itself is printed rounded
    np.testing.assert_allclose(pt_flow.chebyshev_ladder(20, 0.013294),
                               [lib.orc_get_chain_beta(i, 20, 0.013294) for i in range(20)], rtol=1e-15)


def test_step_predictor_manual():
    """steps_i = steps_0 * beta_i^-1/2 * factors; numbers printed at doc/manual.rst:419-440"""
    steps0 = np.array([0.052201, 0.000060, 0.036125, 0.037715])
    factors = np.array([0.887411, 0.887411, 1.044013, 0.887411])
    beta19 = 0.013294
    got = steps0 * beta19 ** -0.5 * factors
    np.testing.assert_allclose(got, [0.401759, 0.000460, 0.327099, 0.290271], rtol=6e-3)


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10"""
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_mt19937_gsl_default_stream():
    """GSL's mt19937 with the default seed 0 (-> 4357): first raw output 4293858116"""
    eng = Oracle("normal", 1, 1, seed=0, rng=RNG_MT19937)
    assert eng.mt_uniform() == 4293858116 / 4294967296.0


def test_evidence_rectangle_rule():
    """reference src/analyse.c:82-93"""
    beta = np.array([1.0, 0.5, 0.1])
    mean_dl = np.array([-10.0, -8.0, -3.0])  # mean of column 2 per chain
    want = (-3.0 / 0.1) * 0.1 + (-8.0 / 0.5) * 0.4 + (-10.0 / 1.0) * 0.5
    assert abs(evidence(beta, mean_dl) - want) < 1e-12


PHASE_FIXTURES = ["c1_phases", "c1_circular_phases", "c1_logistic_phases", "c1_uniform_phases",
                  "c4_phases", "c2_phases", "c1_adapt_phases", "c1_randomswap_phases"]


def _fixture_data(fx):
    if fx["data_file"]:
        return np.loadtxt(os.path.join(GOLDEN, fx["data_file"]))
    return np.array(fx["data"], dtype=float).reshape(-1, fx["n_cols"])


@pytest.mark.parametrize("name", PHASE_FIXTURES)
def test_phases_byte_identical_to_reference(name, tmp_path):
    """calibrate_first -> calibrate_rest -> run in MT19937 mode: calibration_results after
    every phase and every dump file of `run` must equal the reference's byte for byte, and the
    evidence computed from the dumps must print the same."""
    fx = load(name)
    cfg, opts = fx["config"], fx["engine_opts"]
    rows = [tuple(r) for r in fx["rows"]]
    data = _fixture_data(fx)
    wd = str(tmp_path)
    burn = cfg["BURN_IN_ITERATIONS"]

    def engine():
        # every phase of the reference is a fresh process = a fresh MT19937 stream
        e = Oracle(fx["model"], 1, cfg["N_BETA"], n_par=len(rows), seed=cfg["GSL_RNG_SEED"],
                   rng=RNG_MT19937, proposal=opts.get("proposal", 0),
                   circular_mask=opts.get("circular_mask", 0))
        e.set_data(data)
        if opts.get("adapt"):        # -DADAPT
            e.set_adapt(True, opts["adapt"])
        if opts.get("random_swap"):  # -DRANDOMSWAP
            e.set_random_swap(True)
        return e

    def cal_file():
        return open(os.path.join(wd, "calibration_results")).read()

    pt_flow.calibrate_first(engine(), rows, wd, burn_in_iterations=burn)
    assert cal_file() == fx["phases"]["calibrate_first"]
    pt_flow.calibrate_rest(engine(), rows, wd, beta_0=opts.get("beta_0", -0.001), burn_in_iterations=burn)
    assert cal_file() == fx["phases"]["calibrate_rest"]
    tr, n_swap = pt_flow.run(engine(), rows, wd, cfg["MAX_ITERATIONS"])
    assert cal_file() == fx["phases"]["run"]
    for fname, want in fx["dumps"].items():
        path = os.path.join(wd, fname)
        lines = open(path).read().splitlines()
        assert len(lines) == want["n_lines"], fname
        assert lines[:5] == want["head"] and lines[-5:] == want["tail"], fname
        assert hashlib.sha256(open(path, "rb").read()).hexdigest() == want["sha256"], fname
    # analyse_data_probability on the dump text (7 significant digits, "%6e")
    beta = pt_flow.read_calibration_results(os.path.join(wd, "calibration_results"), cfg["N_BETA"], len(rows))[0]
    mean_dl = [np.mean([float(l.split()[1]) for l in open(os.path.join(wd, f"prob-chain{k}.dump"))])
               for k in range(cfg["N_BETA"])]
    assert "%.5f" % evidence(beta, mean_dl) == fx["evidence"]


def test_philox_mode_is_reproducible_and_thread_independent():
    """PHILOX mode: per-chain counter streams make the result independent of the OpenMP
    thread count (the reference's shared-RNG / shared-subiter races are gone)."""
    fx = load("c1_phases")
    rows = [tuple(r) for r in fx["rows"]]
    data = _fixture_data(fx)
    out = []
    for threads in (1, 4):
        e = Oracle("simplesin", 2, 4, seed=5, rng=RNG_PHILOX, n_threads=threads)
        e.set_data(data)
        pt_flow.setup_chains(e, rows)
        e.set_chains(0, 8, beta=np.tile([1.0, 0.7, 0.4, 0.1], 2))
        e.run(6, 25, prob_every=1, params_chains=2)
        out.append((e.read_trace(), e.get_chains()))
    for k in ("prob", "prob_minus_prior", "params"):
        np.testing.assert_array_equal(out[0][0][k], out[1][0][k])
    for k in out[0][1]:
        np.testing.assert_array_equal(out[0][1][k], out[1][1][k])
    assert out[0][1]["n_iter"].tolist() == [150] * 8
    # the two ensembles use different streams
    assert not np.array_equal(out[0][0]["prob"][:, 0], out[0][0]["prob"][:, 4])


def test_stats_fixture_run_is_reproduced_by_the_oracle():
    """The full-configuration statistical fixture (tests/golden/c1_stats.json: 24 `run`s of the
    unmodified reference at C1's real size, N_BETA 20 x 20 000 iterations, from one calibration)
    is pinned like the phase fixtures: the oracle in MT19937 mode, seeded like the fixture's first
    run, reproduces that run's evidence as `analyse` prints it and its posterior means / variances
    of chain 0 (computed from "%.15e" text on the reference side)."""
    fx = load("c1_stats")
    rows = [tuple(r) for r in fx["rows"]]
    n_par, n_beta, iters = len(rows), fx["config"]["N_BETA"], fx["config"]["MAX_ITERATIONS"]
    data = np.loadtxt(os.path.join(GOLDEN, fx["data_file"]))
    cal = np.array(fx["calibration_results"].split(), dtype=float).reshape(n_beta, 1 + 2 * n_par)
    run = fx["runs"][0]
    e = Oracle(fx["model"], 1, n_beta, n_par=n_par, seed=run["seed"], rng=RNG_MT19937, n_threads=1)
    e.set_data(data)
    pt_flow.setup_chains(e, rows)
    pt_flow.apply_calibration(e, 0, cal[:, 0], cal[:, 1:1 + n_par], cal[:, 1 + n_par:])
    n_swap = 2000 // n_beta
    e.run(-(-iters // n_swap), n_swap, prob_every=1, params_chains=1)
    tr = e.read_trace()
    assert tr["params"].shape[0] == run["n"]
    # the reference's analyse reads "%6e" text: 7 significant digits
    dl_text = np.array([[float("%6e" % v) for v in col] for col in tr["prob_minus_prior"].T])
    assert "%.5f" % evidence(cal[:, 0], dl_text.mean(axis=1)) == "%.5f" % run["lnz"]
    p0 = tr["params"][:, 0, :]
    np.testing.assert_allclose(p0.mean(axis=0), run["mean"], rtol=1e-12)
    np.testing.assert_allclose(p0.var(axis=0), run["var"], rtol=1e-9)
