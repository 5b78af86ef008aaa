"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(include/apemost_gpu.h) via ctypes.

Tolerances
  calc_model (fp64): 1e-12 relative against the oracle / the golden fixtures (north_star).
  trajectories: the engine and the oracle (ORC_RNG_PHILOX) share the counter RNG, so chains are
  compared step by step: accept/reject counters must be equal, visited points agree to 1e-9
  relative (device log/cos/sqrt differ from glibc's in the last ulp; the likelihood sum is
  reduced in a different order).
"""
import json
import os

import numpy as np
import pytest

import pt_flow
from oracle_binding import Oracle, RNG_PHILOX

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL_LOGLIK = 1e-12
RTOL_TRAJ = 1e-9


@pytest.fixture(scope="module")
def capi():
    from apemost_b200 import capi as c
    c.load_library()
    return c


def load(name):
    return json.load(open(os.path.join(GOLDEN, name + ".json")))


def lightcurve(n, seed=12345, span=1000.0):
    rng = np.random.default_rng(seed)
    x = np.arange(n) * (span / n)
    y = 1.3 * np.sin(2 * np.pi * 7.25 * x + 0.31 * 2 * np.pi) + 0.2 + rng.normal(0, 0.5, n)
    return np.stack([x, y], axis=1)


SS5_LO, SS5_HI = np.array([0.0, 4.0, 0.0, -1.0]), np.array([3.0, 10.0, 2 * np.pi, 1.0])


# ---------------------------------------------------------------- calc_model
def test_manual_known_answer(capi):
    """reference doc/manual.rst:189-213"""
    e = capi.Engine("simplesin", 1, 1)
    e.set_data(np.array([[101, 0.67], [102, 1.01], [103, 0.79], [104, 1.34]]))
    prob, prior = e.eval([[1, 0.2, 1, 0]])
    assert abs(prob[0] / -1.480898044165363e+01 - 1) < RTOL_LOGLIK
    assert prior[0] == 0.0


@pytest.mark.parametrize("model", ["simplesin", "simplesin5", "simplesin2", "normal", "pulse_vrot", "pulse",
                                   "bernoulli_example"])
def test_calc_model_matches_reference_fixture(capi, model):
    fx = load("eval_" + model)
    n_par = len(fx["rows"])
    data = np.array(fx["data"], dtype=float).reshape(-1, fx["n_cols"])
    params = np.array(fx["params"], dtype=float).reshape(-1, n_par)
    e = capi.Engine(model.replace("_example", ""), 1, 1, n_par=n_par)
    e.set_data(data)
    prob, prior = e.eval(params)
    np.testing.assert_allclose(prob, np.array(fx["prob"], dtype=float), rtol=RTOL_LOGLIK)
    np.testing.assert_allclose(prior, np.array(fx["prior"], dtype=float), rtol=RTOL_LOGLIK)


@pytest.mark.parametrize("n_rows", [1, 7, 2047, 2048, 2049, 6000, 100000])
@pytest.mark.parametrize("model", ["simplesin5", "simplesin"])
def test_calc_model_ragged_sizes(capi, model, n_rows):
    """row counts around the 2048-row chunk: empty tails, exactly full, one over"""
    data = lightcurve(n_rows, seed=n_rows)
    rng = np.random.default_rng(n_rows + 1)
    lo, hi = (SS5_LO, SS5_HI) if model == "simplesin5" else (SS5_LO, np.array([3.0, 10.0, 1.0, 1.0]))
    params = rng.uniform(lo, hi, size=(37, 4))
    beta = rng.uniform(0.01, 1, 37)
    e, o = capi.Engine(model, 1, 1), Oracle(model, 1, 1)
    for eng in (e, o):
        eng.set_data(data)
    p_gpu, _ = e.eval(params, beta)
    p_cpu, _ = o.eval(params, beta)
    np.testing.assert_allclose(p_gpu, p_cpu, rtol=RTOL_LOGLIK)


def test_calc_model_million_rows(capi):
    """config C3's table size: 1M rows, 64 parameter vectors"""
    data = lightcurve(1_000_000)
    rng = np.random.default_rng(3)
    params = rng.uniform(SS5_LO, SS5_HI, size=(64, 4))
    params[0] = [1.3, 7.25, 0.31 * 2 * np.pi, 0.2]
    e, o = capi.Engine("simplesin5", 1, 1), Oracle("simplesin5", 1, 1)
    for eng in (e, o):
        eng.set_data(data)
    p_gpu, _ = e.eval(params)
    p_cpu, _ = o.eval(params)
    np.testing.assert_allclose(p_gpu, p_cpu, rtol=RTOL_LOGLIK)


def test_calc_model_outside_fast_sine_range(capi):
    """|argument| >= 2^30, infinities: the kernel's exact fallback must agree with libm"""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(0, 50, 3000), rng.uniform(1e9, 1e13, 200), [1e300, 3e15]])
    rng.shuffle(x)
    data = np.stack([x, rng.normal(0, 1, x.size)], axis=1)
    params = rng.uniform(SS5_LO, SS5_HI, size=(9, 4))
    e, o = capi.Engine("simplesin5", 1, 1), Oracle("simplesin5", 1, 1)
    for eng in (e, o):
        eng.set_data(data)
    p_gpu, _ = e.eval(params)
    p_cpu, _ = o.eval(params)
    np.testing.assert_allclose(p_gpu, p_cpu, rtol=RTOL_LOGLIK)


def test_wide_table_is_narrowed(capi):
    """a data file with more columns than the model reads (gsl_matrix with tda > 2)"""
    d2 = lightcurve(5000)
    d4 = np.concatenate([d2, np.ones((5000, 2))], axis=1)
    params = np.array([[1.3, 7.25, 1.9, 0.2]])
    e2, e4 = capi.Engine("simplesin5", 1, 1), capi.Engine("simplesin5", 1, 1)
    e2.set_data(d2)
    e4.set_data(d4)
    assert e2.eval(params)[0][0] == e4.eval(params)[0][0]


def test_linearity_in_rows(capi):
    """size-independent property: the running sum over a table equals the sum over its halves"""
    data = lightcurve(300001)
    params = np.array([[1.1, 7.3, 2.0, 0.1], [0.5, 5.0, 0.3, -0.2]])
    full = capi.Engine("simplesin5", 1, 1)
    full.set_data(data)
    a, b = capi.Engine("simplesin5", 1, 1), capi.Engine("simplesin5", 1, 1)
    a.set_data(data[:123457])
    b.set_data(data[123457:])
    np.testing.assert_allclose(full.eval(params)[0], a.eval(params)[0] + b.eval(params)[0], rtol=1e-13)


def test_full_size_properties_c3(capi):
    """BASELINE config C3 at full size (1M rows x 4096 parameter vectors, one launch of the hot
    kernel), through properties that need no CPU evaluation: a vector's value does not depend on
    its slot (which chain tile, which position in the tile), is reproducible bit for bit from
    launch to launch, and a sample of slots agrees with the oracle to 1e-12"""
    data = lightcurve(1_000_000)
    rng = np.random.default_rng(11)
    distinct = rng.uniform(SS5_LO, SS5_HI, size=(64, 4))
    distinct[0] = [1.3, 7.25, 0.31 * 2 * np.pi, 0.2]
    slot_of = rng.integers(0, 64, size=4096)
    params = distinct[slot_of]
    e = capi.Engine("simplesin5", 1, 1)
    e.set_data(data)
    p1, _ = e.eval(params)
    p2, _ = e.eval(params)
    np.testing.assert_array_equal(p1, p2)
    for k in range(64):
        vals = p1[slot_of == k]
        assert (vals == vals[0]).all(), k
    o = Oracle("simplesin5", 1, 1)
    o.set_data(data)
    want, _ = o.eval(distinct[:6])
    got = np.array([p1[slot_of == k][0] for k in range(6)])
    np.testing.assert_allclose(got, want, rtol=RTOL_LOGLIK)


def test_full_size_properties_c5_shard(capi):
    """one GPU's shard of BASELINE config C5 (12.5M rows = 200 MB, more than L2): the running sum
    over the shard equals the sum over two unequal parts (1e-13), and does not depend on the tile
    a vector sits in"""
    n = 12_500_000
    data = lightcurve(n, seed=5)
    params = np.tile(np.array([[1.3, 7.25, 0.31 * 2 * np.pi, 0.2], [0.7, 5.5, 1.0, -0.1]]), (9, 1))  # 18 slots
    full = capi.Engine("simplesin5", 1, 1)
    full.set_data(data)
    p_full, _ = full.eval(params)
    assert (p_full[0::2] == p_full[0]).all() and (p_full[1::2] == p_full[1]).all()
    cut = 4_999_999
    parts = []
    for lo, hi in ((0, cut), (cut, n)):
        e = capi.Engine("simplesin5", 1, 1)
        e.set_data(data[lo:hi])
        parts.append(e.eval(params[:2])[0])
        e.close()
    np.testing.assert_allclose(p_full[:2], parts[0] + parts[1], rtol=1e-13)


# ---------------------------------------------------------------- sampler trajectories
# APM_PATH_TILED / APM_PATH_FUSED / APM_PATH_CLUSTER / APM_PATH_GRID
PATHS = [pytest.param(1, id="tiled"), pytest.param(2, id="fused"), pytest.param(3, id="cluster"),
         pytest.param(4, id="grid")]
# calibration asked for as "cluster" takes the warp-group kernel (group_calibrate_kernel: the selected chains
# dealt out over the SMs, a group of warps per chain); data-free models have the fused kernel only
CALIBRATION_PATH = {1: 1, 2: 2, 3: 3, 4: 4}


def _pair(capi, model, n_ens, n_beta, n_par=None, seed=1, path=0, **kw):
    """the CUDA engine (on the requested kernel path) and the CPU oracle, configured alike"""
    return (capi.Engine(model, n_ens, n_beta, n_par=n_par, seed=seed, path=path, **kw),
            Oracle(model, n_ens, n_beta, n_par=n_par, seed=seed, rng=RNG_PHILOX, **kw))


def _compare_state(st_gpu, st_cpu):
    for k in ("accept", "reject", "params_accepts", "params_rejects", "n_iter", "swapcount", "rng_counter"):
        np.testing.assert_array_equal(st_gpu[k], st_cpu[k], err_msg=k)
    for k in ("params", "params_best", "steps", "prob", "prior", "prob_best", "beta"):
        np.testing.assert_allclose(st_gpu[k], st_cpu[k], rtol=RTOL_TRAJ, atol=1e-300, err_msg=k)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("quirks", [3, 0])
@pytest.mark.parametrize("name,kw", [("c1_phases", {}), ("c1_circular_phases", dict(circular_mask=4)),
                                      ("c1_logistic_phases", dict(proposal=1)),
                                      ("c1_uniform_phases", dict(proposal=2)), ("c4_phases", {}),
                                      ("c2_phases", {})])
def test_run_trajectory_equals_oracle(capi, name, kw, quirks, path):
    """3 ensembles x the fixture's ladder, started from the reference's own calibration_results:
    the whole run (steps, swaps, best tracking, traces, accumulators) against the oracle"""
    fx = load(name)
    if path in (3, 4) and fx["model"] == "normal":
        pytest.skip("the cluster and grid paths are for models with data")
    rows = [tuple(r) for r in fx["rows"]]
    n_par, n_beta, n_ens = len(rows), fx["config"]["N_BETA"], 3
    data = (np.loadtxt(os.path.join(GOLDEN, fx["data_file"])) if fx["data_file"]
            else np.array(fx["data"], dtype=float).reshape(-1, 2))
    cal = np.array(fx["phases"]["calibrate_rest"].split(), dtype=float).reshape(n_beta, 1 + 2 * n_par)
    engines = _pair(capi, fx["model"], n_ens, n_beta, n_par=n_par, seed=17, quirks=quirks, path=path, **kw)
    res = []
    for eng in engines:
        eng.set_data(data)
        pt_flow.setup_chains(eng, rows)
        pt_flow.apply_calibration(eng, 0, np.tile(cal[:, 0], n_ens), np.tile(cal[:, 1:1 + n_par], (n_ens, 1)),
                                  np.tile(cal[:, 1 + n_par:], (n_ens, 1)))
        eng.reset_stats()
        eng.run(8, 40, prob_every=1, params_chains=2)
        res.append((eng.read_trace(), eng.get_chains(), eng.get_stats()))
    (tr_g, st_g, ac_g), (tr_c, st_c, ac_c) = res
    assert engines[0].last_path() == path
    _compare_state(st_g, st_c)
    for k in ("prob", "prob_minus_prior", "params"):
        np.testing.assert_allclose(tr_g[k], tr_c[k], rtol=RTOL_TRAJ, atol=1e-300, err_msg=k)
    np.testing.assert_array_equal(ac_g["n"], ac_c["n"])
    for k in ("sum_dl", "sum_params", "sum_params_sq"):
        np.testing.assert_allclose(ac_g[k], ac_c[k], rtol=RTOL_TRAJ)
    assert st_g["swapcount"].sum() > 0, "the test must exercise accepted swaps"


def test_error_codes_and_the_handle_stays_usable(capi):
    """error behaviour across the C ABI (include/apemost_gpu.h: negative APM_E* codes, a message in
    apm_gpu_last_error, nothing thrown, nothing exits): the reference asserts or exit(1)s in these
    situations; the library reports them and the handle stays usable"""
    E = capi.EngineError

    def code(fn, *a, **kw):
        with pytest.raises(E) as ei:
            fn(*a, **kw)
        assert str(ei.value), "a message must come with the code"
        return ei.value.code

    # create: geometry, parameter count, model id
    assert code(capi.Engine, "simplesin", 0, 4) == -1
    assert code(capi.Engine, "simplesin", 1, 4, n_par=3) == -1
    assert code(capi.Engine, "pulse", 1, 4, n_par=5) == -1
    assert code(capi.Engine, 99, 1, 4, n_par=2) == -1
    e = capi.Engine("simplesin", 2, 5, seed=3)
    data = lightcurve(700)
    # call order: nothing runs before the table and the bounds are there
    assert code(e.eval, np.zeros((1, 4))) == -5
    assert code(e.run, 1, 3) == -5
    assert code(e.calibrate) == -5
    e._marg = (2, 20, 8)                                       # (what set_marginals would have remembered)
    assert code(e.get_marginals) == -5
    # arguments
    assert code(e.set_data, np.zeros((10, 1))) == -1          # the model reads two columns
    assert code(e.set_bounds, [0, 0, 0, 2.0], [1, 1, 1, 1.0]) == -1   # min > max
    assert code(e.set_chains, 8, 5, beta=np.ones(5)) == -1     # chains [8, 13) of 10
    assert code(e.get_chains, 9, 2) == -1
    assert code(e.set_adapt, True, 1.5) == -1
    assert code(e.set_marginals, 7, 20, 5, 8) == -1
    # ... and the handle is as good as new
    e.set_data(data)
    pt_flow.setup_chains(e, [(1.3, 0.0, 3.0, "a", 0.01), (7.25, 4.0, 10.0, "f", 1e-4), (0.31, 0.0, 1.0, "p", 0.01),
                             (0.2, -1.0, 1.0, "o", 0.01)])
    assert code(e.run, 1, 3, prob_every=-1) == -1              # bad trace configuration
    o = Oracle("simplesin", 2, 5, seed=3, rng=RNG_PHILOX)
    o.set_data(data)
    pt_flow.setup_chains(o, [(1.3, 0.0, 3.0, "a", 0.01), (7.25, 4.0, 10.0, "f", 1e-4), (0.31, 0.0, 1.0, "p", 0.01),
                             (0.2, -1.0, 1.0, "o", 0.01)])
    for eng in (e, o):
        eng.set_chains(0, 10, beta=np.tile(np.linspace(1, 0.2, 5), 2))
        eng.run(3, 7)
    _compare_state(e.get_chains(), o.get_chains())
    # a path that does not apply, asked for by name
    f = capi.Engine("simplesin", 1, 4, seed=1, path=2)
    f.set_data(lightcurve(60000))                              # 960 KB: more than an SM's shared memory
    pt_flow.setup_chains(f, [(1.3, 0.0, 3.0, "a", 0.01), (7.25, 4.0, 10.0, "f", 1e-4), (0.31, 0.0, 1.0, "p", 0.01),
                             (0.2, -1.0, 1.0, "o", 0.01)])
    assert code(f.run, 1, 3) == -1


def test_normal_model_value_is_the_oracles_bit_for_bit(capi):
    """apps/normal.c on the device is written for a short chain of dependent instructions (|x - pos| as
    a subtraction with the sign dropped, the divisions by 1..9 as div_by_known, -sigma q^2 / 2 as
    q^2 (-sigma / 2)): every one of these is the reference's value exactly, so calc_model agrees with
    the oracle's plain restatement in every bit, on a dense sweep of the range, at the bumps'
    centres and next to them"""
    rng = np.random.default_rng(5)
    centres = np.exp(np.arange(10.0))
    x = np.concatenate([rng.uniform(-10, 10000, 200_000), rng.uniform(0, 60, 200_000), centres,
                        np.nextafter(centres, 0), np.nextafter(centres, 1e9), (centres + rng.normal(0, 1e-9, (50, 10))).ravel()])
    beta = rng.uniform(0.01, 1.0, x.size)
    e, o = _pair(capi, "normal", 1, 2)
    out = []
    for eng in (e, o):
        eng.set_data(np.zeros((2, 2)))
        eng.set_bounds([-10.0], [10000.0])
        out.append(eng.eval(x[:, None], beta))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("n_ens,n_beta,adapt,kw", [
    pytest.param(1, 3, False, {}, id="1x3"),
    pytest.param(2, 7, True, dict(proposal=1), id="2x7-logistic-adapt"),
    pytest.param(5, 33, False, dict(proposal=2), id="5x33-uniform"),
    pytest.param(2, 90, True, {}, id="2x90-adapt"),
    pytest.param(3, 20, False, dict(circular_mask=1), id="3x20-circular"),
    pytest.param(150, 8, False, {}, id="150x8-more-ensembles-than-sms")])
def test_data_free_kernel_geometries(capi, n_ens, n_beta, adapt, kw):
    """free_run_kernel (a cluster of two CTAs per ensemble: deciders with the chain's state in their
    registers, drawers writing through distributed shared memory, book-keepers) on ladders that do
    not fill the deciding warps, on more ensembles than SMs, with every proposal kind, a circular
    parameter, adapt and wide step widths on the hot rungs (redraws beyond the attempts drawn
    ahead): trajectories, traces and accumulators against the oracle's"""
    rows = [(20.0, 0.0, 60.0, "x", 4.0)]
    n = n_ens * n_beta
    beta = np.tile(np.linspace(1.0, 0.02, n_beta), n_ens)
    res = []
    engines = _pair(capi, "normal", n_ens, n_beta, seed=11, path=2, **kw)
    for eng in engines:
        eng.set_data(np.zeros((2, 2)))
        pt_flow.setup_chains(eng, rows)
        eng.set_chains(0, n, beta=beta, steps=(4.0 * beta ** -0.75)[:, None])
        eng.set_adapt(adapt, 0.5)
        eng.reset_stats()
        eng.run(4, 31, prob_every=1, params_chains=1)
        tr = eng.read_trace()
        eng.run(3, 7)
        res.append((tr, eng.get_chains(), eng.get_stats()))
    (tr_g, st_g, ac_g), (tr_c, st_c, ac_c) = res
    assert engines[0].last_path() == 2
    _compare_state(st_g, st_c)
    for k in ("prob", "prob_minus_prior", "params"):
        np.testing.assert_allclose(tr_g[k], tr_c[k], rtol=RTOL_TRAJ, atol=1e-300, err_msg=k)
    np.testing.assert_array_equal(ac_g["n"], ac_c["n"])
    for k in ("sum_dl", "sum_params", "sum_params_sq"):
        np.testing.assert_allclose(ac_g[k], ac_c[k], rtol=RTOL_TRAJ)
    assert st_g["accept"].sum() > 0 and st_g["reject"].sum() > 0


def test_data_free_long_run_equals_oracle(capi):
    """31 000 Metropolis steps of 2 x 64 chains of apps/normal.c in ONE launch of free_run_kernel
    (2000 batches, 1000 swap rounds) with -DADAPT on and the counters started next to its
    thresholds: ensemble 0 adapts from the first round and has its counters reset at 100 000,
    ensemble 1 starts adapting at 20 000; the hot rungs' step widths grow to several times the
    range, so most of their proposals need more redraws than are drawn ahead.  State, counters
    and accumulators equal the oracle's at the end: every one of the 4 M accept decisions and
    23 000 adapt decisions is the oracle's.  (Not longer: with a flat hot rung the reference's adapt rule grows
    the step width without bound, and the redraw loop with it -- in the reference as here.)"""
    rows = [(20.0, -10.0, 10000.0, "x", 4.0)]
    n_ens, n_beta = 2, 64
    n = n_ens * n_beta
    beta = np.tile(np.linspace(1.0, 0.01, n_beta), n_ens)
    start = 30.0 * beta ** -0.5
    pa = np.full((n, 1), 52000, dtype=np.uint64)
    pa[n_beta:] = 2000
    res = []
    engines = _pair(capi, "normal", n_ens, n_beta, seed=23, path=2)
    for eng in engines:
        eng.set_data(np.zeros((2, 2)))
        pt_flow.setup_chains(eng, rows)
        eng.set_chains(0, n, beta=beta, steps=start[:, None], params_accepts=pa,
                       params_rejects=(pa * 0.9).astype(np.uint64))
        eng.set_adapt(True, 0.5)
        eng.reset_stats()
        eng.run(1000, 31)
        res.append((eng.get_chains(), eng.get_stats()))
    (st_g, ac_g), (st_c, ac_c) = res
    assert engines[0].last_path() == 2
    _compare_state(st_g, st_c)   # counters equal; values to 1e-9 (the jumps' log / sqrt / cos are the device library's)
    np.testing.assert_array_equal(st_g["steps"], st_c["steps"])
    np.testing.assert_array_equal(ac_g["n"], ac_c["n"])
    for k in ("sum_dl", "sum_params", "sum_params_sq"):
        np.testing.assert_allclose(ac_g[k], ac_c[k], rtol=RTOL_TRAJ)
    assert (st_g["n_iter"] == 31000).all()
    assert (st_g["params_accepts"] + st_g["params_rejects"] < 100000).all(), "adapt must have reset ensemble 0's counters"
    ratio = st_g["steps"][:, 0] / start
    assert ratio.max() > 50 and ratio.min() < 0.05, "adapt must have moved the step widths a long way"


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name,which", [("c1_phases", 2), ("c4_phases", 1), ("c2_phases", 2)])
def test_marginal_statistics_equal_oracle(capi, name, which, path):
    """apm_gpu_set_marginals: the per-parameter histograms (gsl_histogram_increment's
    bins, create_hist's edges) and calc_mcmc_error's batch means, accumulated in the kernels of every
    path, against the oracle's: counts equal, batch means to 1e-9, over two run calls"""
    fx = load(name)
    if path in (3, 4) and fx["model"] == "normal":
        pytest.skip("the cluster and grid paths are for models with data")
    rows = [tuple(r) for r in fx["rows"]]
    n_par, n_beta, n_ens = len(rows), fx["config"]["N_BETA"], 3
    data = (np.loadtxt(os.path.join(GOLDEN, fx["data_file"])) if fx["data_file"]
            else np.array(fx["data"], dtype=float).reshape(-1, 2))
    cal = np.array(fx["phases"]["calibrate_rest"].split(), dtype=float).reshape(n_beta, 1 + 2 * n_par)
    res = []
    for eng in _pair(capi, fx["model"], n_ens, n_beta, n_par=n_par, seed=41, path=path):
        eng.set_data(data)
        pt_flow.setup_chains(eng, rows)
        pt_flow.apply_calibration(eng, 0, np.tile(cal[:, 0], n_ens), np.tile(cal[:, 1:1 + n_par], (n_ens, 1)),
                                  np.tile(cal[:, 1 + n_par:], (n_ens, 1)))
        eng.set_marginals(which, n_bins=50, batch_size=7, max_batches=40)
        eng.run(4, 30)
        eng.run(3, 30, prob_every=1, params_chains=1)
        res.append((eng.get_marginals(), eng.get_chains()))
    (m_g, st_g), (m_c, st_c) = res
    _compare_state(st_g, st_c)
    assert (m_g["n_values"] == 210).all() and (m_g["n_batches"] == 30).all()
    np.testing.assert_array_equal(m_g["n_values"], m_c["n_values"])
    np.testing.assert_array_equal(m_g["n_batches"], m_c["n_batches"])
    np.testing.assert_array_equal(m_g["counts"], m_c["counts"])
    assert (m_g["counts"].sum(axis=2) == 210).all()
    np.testing.assert_allclose(m_g["batch_means"][:, :, :30], m_c["batch_means"][:, :, :30], rtol=RTOL_TRAJ, atol=1e-300)


def _logistic_table(n_rows, n_cols, seed):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n_rows, n_cols - 1))
    w = np.array([0.5, -0.3, 0.8])[:n_cols - 1]
    y = (rng.uniform(size=n_rows) < 1 / (1 + np.exp(-(0.1 + X @ w)))).astype(float)
    return np.column_stack([y, X])


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("n_cols,n_rows", [(3, 700), (2, 333), (4, 5000)])
def test_bernoulli_rows_of_four_doubles(capi, path, n_cols, n_rows):
    """apps/bernoulli_example.c: a model reading 2..4 data columns (n_par = columns), i.e. table
    rows of 4 doubles on the device; calc_model to 1e-12 and whole runs against the oracle on
    every kernel path (5000 rows x 32 B does not fit the fused paths' shared memory twice over,
    so that case is tiled only)"""
    if n_rows * 32 > 150_000 and path in (2, 3):
        pytest.skip("table too large for the one-SM shared-memory paths")
    data = _logistic_table(n_rows, n_cols, 7 * n_cols)
    n_ens, n_beta, n_par = 2, 6, n_cols
    n = n_ens * n_beta
    rng = np.random.default_rng(n_cols)
    lo, hi = np.full(n_par, -5.0), np.full(n_par, 5.0)
    params = rng.normal(0, 0.3, (n, n_par))
    beta = np.tile(np.linspace(1.0, 0.5, n_beta), n_ens)
    steps = np.full((n, n_par), 0.15) * beta[:, None] ** -0.5
    res = []
    for eng in _pair(capi, "bernoulli", n_ens, n_beta, n_par=n_par, seed=53, path=path):
        eng.set_data(data)
        eng.set_bounds(lo, hi)
        prob, prior = eng.eval(params, beta)
        eng.set_chains(0, n, beta=beta, params=params, steps=steps, params_best=params, prob=prob, prior=prior)
        eng.run(5, 20, prob_every=1, params_chains=1)
        res.append((prob, prior, eng.read_trace(), eng.get_chains()))
    (p_g, pr_g, tr_g, st_g), (p_c, pr_c, tr_c, st_c) = res
    np.testing.assert_allclose(p_g, p_c, rtol=RTOL_LOGLIK)
    np.testing.assert_allclose(pr_g, pr_c, rtol=RTOL_LOGLIK, atol=1e-300)
    _compare_state(st_g, st_c)
    np.testing.assert_allclose(tr_g["prob"], tr_c["prob"], rtol=RTOL_TRAJ)
    assert st_g["swapcount"].sum() > 0 and 0.05 < st_g["accept"].sum() / (n * 100) < 0.95


def _pulse_spectrum(n_bins, modes, lifetime, seed):
    rng = np.random.default_rng(seed)
    nu = np.linspace(90.0, 110.0, n_bins)
    y = sum(h / (1 + (2 * np.pi * (f - nu) * lifetime) ** 2) for f, h in modes)
    return np.stack([nu, y * rng.exponential(1.0, n_bins)], axis=1)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("n_modes", [1, 2, 3, 5, 7])
def test_pulse_any_number_of_modes(capi, path, n_modes):
    """apps/pulse.c:12-56 takes n_par = 2 + 2k for k modes.  The device row term puts the k Lorentzians
    over one denominator (ModelPulse::accum_fast, one division per bin instead of k + 1): calc_model
    to 1e-12 and whole runs against the oracle's sum of quotients for k = 1 .. 7 on every kernel path;
    a non-positive height (the common-denominator form is not used then) takes the reference-order
    fallback and must agree as well"""
    modes = [(92.0 + 2.4 * j, 5.0 - 0.5 * j) for j in range(n_modes)]
    data = _pulse_spectrum(600, modes, 0.5, 11 * n_modes)
    n_ens, n_beta, n_par = 2, 6, 2 + 2 * n_modes
    n = n_ens * n_beta
    truth = np.array([0.5, 0.0] + [v for m in modes for v in m])
    lo = np.array([0.05, -1.0] + [v for f, h in modes for v in (f - 1.0, 0.0)])
    hi = np.array([2.0, 1.0] + [v for f, h in modes for v in (f + 1.0, 20.0)])
    rng = np.random.default_rng(n_modes)
    params = np.clip(truth * (1 + rng.normal(0, 0.02, (n, n_par))), lo, hi)
    params[:, 1] = rng.normal(0, 0.1, n)
    beta = np.tile(np.linspace(1.0, 0.4, n_beta), n_ens)
    steps = np.tile(0.01 * (hi - lo), (n, 1)) * beta[:, None] ** -0.5
    edge = params[:4].copy()
    edge[0, 3] = 0.0       # h_1 = 0: y loses a term
    edge[1, 3] = -1e-3     # a negative height
    edge[2, 0] = 1e12      # very long lifetime: d_j ~ 1e27
    edge[3, 3] = 1e35      # beyond fast_ok's height bound
    res = []
    for eng in _pair(capi, "pulse", n_ens, n_beta, n_par=n_par, seed=71, path=path):
        eng.set_data(data)
        eng.set_bounds(lo, hi)
        e_prob, e_prior = eng.eval(edge, beta[:4])
        prob, prior = eng.eval(params, beta)
        eng.set_chains(0, n, beta=beta, params=params, steps=steps, params_best=params, prob=prob, prior=prior)
        eng.run(5, 20, prob_every=1, params_chains=1)
        res.append((prob, prior, e_prob, e_prior, eng.read_trace(), eng.get_chains()))
    (p_g, pr_g, ep_g, epr_g, tr_g, st_g), (p_c, pr_c, ep_c, epr_c, tr_c, st_c) = res
    np.testing.assert_allclose(p_g, p_c, rtol=RTOL_LOGLIK)
    np.testing.assert_allclose(pr_g, pr_c, rtol=RTOL_LOGLIK, atol=1e-300)
    np.testing.assert_allclose(ep_g, ep_c, rtol=RTOL_LOGLIK, equal_nan=True)
    np.testing.assert_allclose(epr_g, epr_c, rtol=RTOL_LOGLIK, equal_nan=True)
    _compare_state(st_g, st_c)
    np.testing.assert_allclose(tr_g["prob"], tr_c["prob"], rtol=RTOL_TRAJ)
    assert 0.02 < st_g["accept"].sum() / (n * 100) < 0.98


@pytest.mark.parametrize("path", PATHS)
def test_adapt_trajectory_equals_oracle(capi, path):
    """-DADAPT (reference src/parallel_tempering.c:282-302): once the counter sums reach 20000 the
    step widths move by 1 % per round, beyond 100000 the counters are reset; the oracle's adapt is
    byte-pinned to the reference build by tests/golden/c1_adapt_phases.json"""
    fx = load("c1_adapt_phases")
    rows = [tuple(r) for r in fx["rows"]]
    n_par, n_beta, n_ens = len(rows), fx["config"]["N_BETA"], 2
    data = np.loadtxt(os.path.join(GOLDEN, fx["data_file"]))
    cal = np.array(fx["phases"]["calibrate_rest"].split(), dtype=float).reshape(n_beta, 1 + 2 * n_par)
    res = []
    for eng in _pair(capi, fx["model"], n_ens, n_beta, n_par=n_par, seed=29, path=path):
        eng.set_data(data)
        pt_flow.setup_chains(eng, rows)
        pt_flow.apply_calibration(eng, 0, np.tile(cal[:, 0], n_ens), np.tile(cal[:, 1:1 + n_par], (n_ens, 1)),
                                  np.tile(cal[:, 1 + n_par:], (n_ens, 1)))
        # start the counters just below the two thresholds so that a short run crosses both
        n = n_ens * n_beta
        pa = np.full((n, n_par), 2600, dtype=np.uint64)
        pa[n_beta:] = 13000
        eng.set_chains(0, n, params_accepts=pa, params_rejects=(pa * 0.9).astype(np.uint64))
        eng.set_adapt(True, 0.5)
        eng.run(12, 50, prob_every=1, params_chains=1)
        res.append((eng.read_trace(), eng.get_chains()))
    (tr_g, st_g), (tr_c, st_c) = res
    _compare_state(st_g, st_c)
    np.testing.assert_allclose(tr_g["prob"], tr_c["prob"], rtol=RTOL_TRAJ)
    assert not np.allclose(st_g["steps"][:n_beta], cal[:, 1:1 + n_par]), "adapt must have rescaled the steps"
    assert (st_g["params_accepts"][n_beta:] < 13000).all(), "the counters of ensemble 1 must have been reset"


@pytest.mark.parametrize("path", PATHS)
def test_run_continues_across_calls(capi, path):
    """two runs of 4 rounds == one run of 8 rounds (state, RNG position and swap stream persist)"""
    fx = load("c1_phases")
    rows = [tuple(r) for r in fx["rows"]]
    data = np.loadtxt(os.path.join(GOLDEN, fx["data_file"]))
    outs = []
    for split in (False, True):
        e = capi.Engine("simplesin", 2, 4, seed=3, path=path)
        e.set_data(data)
        pt_flow.setup_chains(e, rows)
        e.set_chains(0, 8, beta=np.tile([1.0, 0.6, 0.3, 0.1], 2))
        if split:
            e.run(4, 25)
            e.run(4, 25)
        else:
            e.run(8, 25)
        outs.append(e.get_chains())
    for k in outs[0]:
        np.testing.assert_array_equal(outs[0][k], outs[1][k], err_msg=k)


@pytest.mark.parametrize("path", PATHS)
def test_paths_are_bit_reproducible(capi, path):
    """the same run six times over: every array must come back bit for bit the same (the paths with
    warp groups, clusters and grid barriers have plenty of places where a missing synchronisation
    would show up as run-to-run differences; compute-sanitizer is not available on this pool)"""
    n_rows = 1522 if path in (2, 3) else 40_000
    data = lightcurve(n_rows, seed=3)
    n_ens, n_beta = 3, 20
    n = n_ens * n_beta
    rng = np.random.default_rng(8)
    scale = (1e6 / n_rows) ** 0.5
    params = np.tile([1.3, 7.25, 0.31 * 2 * np.pi, 0.2], (n, 1)) + rng.normal(0, 1e-5, (n, 4)) * scale
    beta = np.tile(np.linspace(1.0, 0.7, n_beta), n_ens)
    steps = np.tile([7e-4, 3e-7, 1.1e-3, 5e-4], (n, 1)) * scale * beta[:, None] ** -0.5
    outs = []
    for rep in range(6):
        e = capi.Engine("simplesin5", n_ens, n_beta, seed=71, path=path)
        e.set_data(data)
        e.set_bounds(SS5_LO, SS5_HI)
        e.set_chains(0, n, beta=beta, params=params, steps=steps, params_best=params)
        e.set_adapt(True, 0.5)
        e.reset_stats()
        e.run(7, 13, prob_every=1, params_chains=2)
        outs.append((e.get_chains(), e.read_trace(), e.get_stats()))
        e.close()
    for st, tr, ac in outs[1:]:
        for ref, got in ((outs[0][0], st), (outs[0][1], tr), (outs[0][2], ac)):
            for k in ref:
                np.testing.assert_array_equal(ref[k], got[k], err_msg=k)


def test_single_chain_ladder_never_swaps(capi):
    data = lightcurve(3000)
    for eng in _pair(capi, "simplesin5", 5, 1, seed=2):
        eng.set_data(data)
        eng.set_bounds(SS5_LO, SS5_HI)
        eng.set_chains(0, 5, params=np.tile([1.3, 7.25, 1.9, 0.2], (5, 1)), steps=np.tile([0.01, 1e-5, 0.01, 0.01], (5, 1)))
        eng.run(3, 10)
        assert eng.get_chains()["swapcount"].sum() == 0


@pytest.mark.parametrize("n_ens,n_beta,expect", [(1, 20, "8 CTAs x (2..3 chains x 5 warps)"),
                                                 (1, 64, "8 CTAs x (8 chains x 2 warps)"),
                                                 (2, 40, "8 CTAs x (5 chains x 3 warps)"),
                                                 (64, 20, "2 CTAs x (10 chains x 1 warp)"),
                                                 (3, 2, "2 CTAs x (1 chain x 16 warps)"),
                                                 (5, 7, "4 CTAs x (1..2 chains x 8 warps)")])
def test_cluster_path_geometries(capi, n_ens, n_beta, expect):
    """the cluster path (ensemble spread over 2..8 CTAs, warp groups per chain, swaps through
    distributed shared memory) against the oracle, for ladders that split evenly and unevenly"""
    data = lightcurve(1522, seed=n_beta)
    n = n_ens * n_beta
    rng = np.random.default_rng(n)
    params = np.tile([1.3, 7.25, 0.31 * 2 * np.pi, 0.2], (n, 1)) + rng.normal(0, 1e-4, (n, 4))
    beta = np.tile(np.linspace(1.0, 0.7, n_beta), n_ens)   # flat ladder: swaps are accepted often
    steps = np.tile([2e-2, 3e-5, 2e-2, 1e-2], (n, 1)) * beta[:, None] ** -0.5
    res = []
    for eng in _pair(capi, "simplesin5", n_ens, n_beta, seed=41, path=3):
        eng.set_data(data)
        eng.set_bounds(SS5_LO, SS5_HI)
        eng.set_chains(0, n, beta=beta, params=params, steps=steps, params_best=params)
        eng.reset_stats()
        eng.run(6, 9, prob_every=1, params_chains=1)
        eng.run(5, 4, prob_every=2, params_chains=2)
        res.append((eng.read_trace(), eng.get_chains(), eng.get_stats()))
    (tr_g, st_g, ac_g), (tr_c, st_c, ac_c) = res
    _compare_state(st_g, st_c)
    for k in ("prob", "prob_minus_prior", "params"):
        np.testing.assert_allclose(tr_g[k], tr_c[k], rtol=RTOL_TRAJ, atol=1e-300, err_msg=k)
    np.testing.assert_array_equal(ac_g["n"], ac_c["n"])
    np.testing.assert_allclose(ac_g["sum_dl"], ac_c["sum_dl"], rtol=RTOL_TRAJ)
    assert st_g["swapcount"].sum() > 0, expect


def test_auto_path_prefers_cluster_for_few_ensembles(capi):
    data = lightcurve(1522)
    for n_ens, want in ((1, 3), (64, 3), (100, 2)):
        e = capi.Engine("simplesin5", n_ens, 20, seed=1)
        e.set_data(data)
        e.set_bounds(SS5_LO, SS5_HI)
        n = n_ens * 20
        e.set_chains(0, n, params=np.tile([1.3, 7.25, 1.9, 0.2], (n, 1)), steps=np.tile([0.01, 1e-5, 0.01, 0.01], (n, 1)),
                     beta=np.tile(np.linspace(1, 0.1, 20), n_ens))
        e.run(1, 3)
        assert e.last_path() == want, (n_ens, e.last_path())


@pytest.mark.parametrize("n_rows,n_ens,n_beta", [(40_000, 1, 20), (300_001, 3, 7), (100, 2, 5), (1_200_000, 1, 4)])
def test_grid_path_mid_size_tables(capi, n_rows, n_ens, n_beta):
    """the grid path (table partitioned over the shared memories of all SMs, one grid barrier per
    step, decisions replicated in every CTA) against the oracle: tables from fewer rows than SMs to
    1.2 M rows, with traces, accumulators and accepted swaps"""
    data = lightcurve(n_rows, seed=n_beta)
    n = n_ens * n_beta
    rng = np.random.default_rng(n)
    scale = (1e6 / n_rows) ** 0.5
    params = np.tile([1.3, 7.25, 0.31 * 2 * np.pi, 0.2], (n, 1)) + rng.normal(0, 1e-5, (n, 4)) * scale
    beta = np.tile(np.linspace(1.0, 0.8, n_beta), n_ens)
    steps = np.tile([7e-4, 3e-7, 1.1e-3, 5e-4], (n, 1)) * scale * beta[:, None] ** -0.5
    res = []
    for eng in _pair(capi, "simplesin5", n_ens, n_beta, seed=61, path=4):
        eng.set_data(data)
        eng.set_bounds(SS5_LO, SS5_HI)
        eng.set_chains(0, n, beta=beta, params=params, steps=steps, params_best=params)
        eng.reset_stats()
        rounds, n_swap = (2, 4) if n_rows > 1_000_000 else (6, 9)
        eng.run(rounds, n_swap, prob_every=1, params_chains=1)
        eng.run(2, 3, prob_every=2, params_chains=2)
        res.append((eng.read_trace(), eng.get_chains(), eng.get_stats()))
    (tr_g, st_g, ac_g), (tr_c, st_c, ac_c) = res
    _compare_state(st_g, st_c)
    for k in ("prob", "prob_minus_prior", "params"):
        np.testing.assert_allclose(tr_g[k], tr_c[k], rtol=RTOL_TRAJ, atol=1e-300, err_msg=k)
    np.testing.assert_array_equal(ac_g["n"], ac_c["n"])
    np.testing.assert_allclose(ac_g["sum_dl"], ac_c["sum_dl"], rtol=RTOL_TRAJ)


def test_fused_path_refuses_a_table_that_does_not_fit(capi):
    e = capi.Engine("simplesin5", 1, 2, path=2)
    e.set_data(lightcurve(20_000))
    e.set_bounds(SS5_LO, SS5_HI)
    with pytest.raises(Exception, match="shared memory"):
        e.run(1, 1)


def test_fused_ladder_longer_than_the_cta(capi):
    """40 rungs: warps take several chains each (FUSED_MAX_WARPS = 16); fused == tiled == oracle"""
    data = lightcurve(1500)
    n_ens, n_beta = 2, 40
    n = n_ens * n_beta
    rng = np.random.default_rng(5)
    params = np.tile([1.3, 7.25, 0.31 * 2 * np.pi, 0.2], (n, 1)) + rng.normal(0, 1e-4, (n, 4))
    beta = np.tile(pt_flow.chebyshev_ladder(n_beta, 0.01), n_ens)
    steps = np.tile([2e-2, 3e-5, 2e-2, 1e-2], (n, 1)) * beta[:, None] ** -0.5
    states = []
    for path in (1, 2):
        for eng in _pair(capi, "simplesin5", n_ens, n_beta, seed=31, path=path)[:1 if path == 2 else 2]:
            eng.set_data(data)
            eng.set_bounds(SS5_LO, SS5_HI)
            eng.set_chains(0, n, beta=beta, params=params, steps=steps, params_best=params)
            eng.run(5, 12, prob_every=1, params_chains=1)
            states.append(eng.get_chains())
    _compare_state(states[0], states[1])
    _compare_state(states[2], states[1])
    assert states[2]["swapcount"].sum() > 0


def test_big_table_run_trajectory(capi):
    """a table that spans many chunks and row splits (300k rows), 2 x 8 chains"""
    data = lightcurve(300_000)
    n_ens, n_beta = 2, 8
    n = n_ens * n_beta
    rng = np.random.default_rng(0)
    params = np.tile([1.3, 7.25, 0.31 * 2 * np.pi, 0.2], (n, 1)) + rng.normal(0, 1e-5, (n, 4))
    beta = np.tile(pt_flow.chebyshev_ladder(n_beta, 0.01), n_ens)
    steps = np.tile([3e-3, 2e-6, 3e-3, 2e-3], (n, 1)) * beta[:, None] ** -0.5
    res = []
    for eng in _pair(capi, "simplesin5", n_ens, n_beta, seed=9):
        eng.set_data(data)
        eng.set_bounds(SS5_LO, SS5_HI)
        eng.set_chains(0, n, beta=beta, params=params, steps=steps, params_best=params)
        eng.run(2, 10, prob_every=1, params_chains=1)
        res.append((eng.read_trace(), eng.get_chains()))
    _compare_state(res[0][1], res[1][1])
    np.testing.assert_allclose(res[0][0]["prob"], res[1][0]["prob"], rtol=RTOL_TRAJ)


# ---------------------------------------------------------------- calibration
@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", ["c1_phases", "c4_phases", "c2_phases"])
def test_calibration_trajectory_equals_oracle(capi, name, path):
    """markov_chain_calibrate (burn_in + _orig) for every chain of 2 ensembles concurrently:
    final step widths, positions, counters and the progress rows against the oracle"""
    fx = load(name)
    rows = [tuple(r) for r in fx["rows"]]
    n_par, n_beta, n_ens = len(rows), fx["config"]["N_BETA"], 2
    data = (np.loadtxt(os.path.join(GOLDEN, fx["data_file"])) if fx["data_file"]
            else np.array(fx["data"], dtype=float).reshape(-1, 2))
    cal = np.array(fx["phases"]["calibrate_rest"].split(), dtype=float).reshape(n_beta, 1 + 2 * n_par)
    res = []
    engines = _pair(capi, fx["model"], n_ens, n_beta, n_par=n_par, seed=23, path=path)
    for eng in engines:
        eng.set_data(data)
        start, lo, hi, names, step = pt_flow.setup_chains(eng, rows)
        n = n_ens * n_beta
        beta = np.tile(cal[:, 0], n_ens)
        eng.set_chains(0, n, beta=beta, steps=np.tile(step, (n, 1)) * beta[:, None] ** -0.5)
        st = eng.get_chains(fields=("params", "beta"))
        prob, prior = eng.eval(st["params"], st["beta"])
        eng.set_chains(0, n, prob=prob, prior=prior)
        status, prog = eng.calibrate(burn_in_iterations=600, progress_capacity=100000)
        res.append((status, prog, eng.get_chains()))
    (s_g, p_g, st_g), (s_c, p_c, st_c) = res
    assert engines[0].last_path() == (2 if (path in (3, 4) and fx["model"] == "normal") else CALIBRATION_PATH[path])
    np.testing.assert_array_equal(s_g, s_c)
    assert (s_g == 0).all()
    _compare_state(st_g, st_c)
    assert len(p_g) == len(p_c) and len(p_g) > 0
    key = lambda r: (r[0], r[2], r[1])
    for a, b in zip(sorted(p_g, key=key), sorted(p_c, key=key)):
        assert a[:3] == b[:3]
        np.testing.assert_allclose(a[3:], b[3:], rtol=RTOL_TRAJ)


@pytest.mark.parametrize("n_ens,n_beta,pick,ng", [(1, 20, [0], 1), (1, 20, [1], 1), (1, 20, None, 1), (45, 4, None, 2),
                                                   (3, 7, [0, 7, 14], 1)])
def test_group_calibration_geometries(capi, n_ens, n_beta, pick, ng):
    """the warp-group calibration kernel (a few selected chains dealt out over the SMs, 16 / ng warps
    per chain): calibrate_first's shape (ONE chain of a 20-rung ladder), calibrate_rest's (chain 1),
    a whole ladder (a chain per CTA), more chains than SMs (two per CTA, one after the other per
    group) and chain 0 of several ensembles -- final state and progress rows against the oracle"""
    fx = load("c1_phases")
    rows = [tuple(r) for r in fx["rows"]]
    n_par = len(rows)
    data = np.loadtxt(os.path.join(GOLDEN, fx["data_file"]))
    n = n_ens * n_beta
    sel = None
    if pick is not None:
        sel = np.zeros(n, dtype=np.uint8)
        sel[pick] = 1
    res = []
    engines = _pair(capi, fx["model"], n_ens, n_beta, n_par=n_par, seed=29)
    for eng in engines:
        eng.set_data(data)
        start, lo, hi, names, step = pt_flow.setup_chains(eng, rows)
        beta = np.tile(pt_flow.chebyshev_ladder(n_beta, 0.02), n_ens)
        eng.set_chains(0, n, beta=beta, steps=np.tile(step, (n, 1)) * beta[:, None] ** -0.5)
        st = eng.get_chains(fields=("params", "beta"))
        prob, prior = eng.eval(st["params"], st["beta"])
        eng.set_chains(0, n, prob=prob, prior=prior)
        status, prog = eng.calibrate(select=sel, burn_in_iterations=400, iter_readjust=40, no_rescaling_limit=3,
                                     progress_capacity=200000, raise_on_failure=False)
        res.append((status, prog, eng.get_chains()))
    (s_g, p_g, st_g), (s_c, p_c, st_c) = res
    assert engines[0].last_path() == 3
    np.testing.assert_array_equal(s_g, s_c)
    if sel is not None:
        assert (s_g[sel == 0] == -1).all()
    _compare_state(st_g, st_c)
    assert len(p_g) == len(p_c) and len(p_g) > 0
    key = lambda r: (r[0], r[2], r[1])
    for a, b in zip(sorted(p_g, key=key), sorted(p_c, key=key)):
        assert a[:3] == b[:3]
        np.testing.assert_allclose(a[3:], b[3:], rtol=RTOL_TRAJ)


@pytest.mark.parametrize("path", [pytest.param(1, id="tiled"), pytest.param(2, id="fused")])
def test_redraws_beyond_a_million_attempts(capi, path):
    """a step width a million times the parameter's range: the truncated proposal (reference
    src/markov_chain.c:235-240, redraw until inside the bounds) needs ~2.5e6 redraws per coordinate;
    the attempt counter has more than its 20 low bits, so the stream does not repeat
    and the loop ends, on the device as in the oracle, with the same point"""
    rows = [(100.0, -10.0, 10000.0, "x", -1.0)]
    res = []
    for eng in _pair(capi, "normal", 1, 3, seed=5, path=path):
        eng.set_data(np.zeros((2, 2)))
        pt_flow.setup_chains(eng, rows)
        eng.set_chains(0, 3, beta=[1.0, 0.5, 0.25], steps=np.full((3, 1), 1e6 * 10010.0))
        eng.run(1, 2)
        res.append(eng.get_chains())
    _compare_state(res[0], res[1])
    assert (res[0]["params"] >= -10).all() and (res[0]["params"] <= 10000).all()


@pytest.mark.parametrize("path", PATHS)
def test_calibration_selection_and_failure_status(capi, path):
    """only selected chains move; a chain whose acceptance rate cannot be brought down (flat
    likelihood) ends with the reference's 'iteration limit' failure (markov_chain_calibrate.c
    :1169-1174) reported per chain instead of exit(1)"""
    rows = [(100.0, -10.0, 10000.0, "x", -1.0)]
    res = []
    for eng in _pair(capi, "normal", 1, 3, seed=4, path=path):
        eng.set_data(np.zeros((2, 2)))
        pt_flow.setup_chains(eng, rows)
        eng.set_chains(0, 3, beta=[1.0, 0.5, 0.001])
        sel = np.array([0, 0, 1], dtype=np.uint8)
        status, _ = eng.calibrate(select=sel, burn_in_iterations=400, iter_limit=3000, raise_on_failure=False)
        res.append((status, eng.get_chains()))
    np.testing.assert_array_equal(res[0][0], res[1][0])
    assert res[0][0].tolist() == [-1, -1, 2]
    assert res[0][1]["rng_counter"][:2].tolist() == [0, 0]
    _compare_state(res[0][1], res[1][1])


@pytest.mark.parametrize("path", PATHS)
def test_steps_accept_log_equals_oracle(capi, path):
    """apm_gpu_steps: n x { markov_chain_step_for / markov_chain_step ; mcmc_check_best } with the accept
    log assess_acceptance_rate keeps (reference src/markov_chain.c:143-172): the log is equal bit for
    bit, the chains' state to 1e-9, for single-parameter and full steps, on a selection of chains"""
    fx = load("c1_phases")
    rows = [tuple(r) for r in fx["rows"]]
    n_par, n_beta, n_ens = len(rows), fx["config"]["N_BETA"], 2
    data = np.loadtxt(os.path.join(GOLDEN, fx["data_file"]))
    cal = np.array(fx["phases"]["calibrate_rest"].split(), dtype=float).reshape(n_beta, 1 + 2 * n_par)
    res = []
    for eng in _pair(capi, fx["model"], n_ens, n_beta, n_par=n_par, seed=37, path=path):
        eng.set_data(data)
        pt_flow.setup_chains(eng, rows)
        pt_flow.apply_calibration(eng, 0, np.tile(cal[:, 0], n_ens), np.tile(cal[:, 1:1 + n_par], (n_ens, 1)),
                                  np.tile(cal[:, 1 + n_par:], (n_ens, 1)))
        st = eng.get_chains(fields=("params", "beta"))
        prob, prior = eng.eval(st["params"], st["beta"])
        eng.set_chains(0, n_ens * n_beta, prob=prob, prior=prior)
        sel = np.zeros(n_ens * n_beta, dtype=np.uint8)
        sel[[0, 3, 5]] = 1
        logs = [eng.steps(1, 40, select=sel), eng.steps(n_par, 333, select=sel), eng.steps(0, 7)]
        res.append((logs, eng.get_chains()))
    (lg, st_g), (lc, st_c) = res
    for a, b in zip(lg, lc):
        np.testing.assert_array_equal(a, b)
    assert lg[0][:, [1, 2, 4, 6, 7]].sum() == 0 and 0 < lg[1][:, 0].mean() < 1
    _compare_state(st_g, st_c)


def test_host_uniform_stream_equals_oracle(capi):
    """apm_gpu_host_uniform: the chain's host-side stream (counter RNG, purpose 4) is the oracle's,
    number for number, is a stream of its own per chain, and leaves the step streams alone"""
    e, o = _pair(capi, "normal", 2, 3, seed=77)
    got = [[eng.host_uniform(g) for g in (0, 4, 0, 0, 5, 4)] for eng in (e, o)]
    assert got[0] == got[1]
    assert len(set(got[0])) == 6 and all(0.0 < u < 1.0 for u in got[0])
    assert e.get_chains(fields=("rng_counter",))["rng_counter"].sum() == 0


# ---------------------------------------------------------------- statistical parity with the reference
@pytest.mark.parametrize("name", ["c1_phases", "c4_phases"])
def test_evidence_and_posterior_match_reference_statistics(capi, name):
    """Second half of 'correctness' (north_star): with different RNG streams the engine's
    posterior moments and thermodynamic-integration evidence must agree with the reference's.
    Protocol: same params file, same calibration_results (the reference's own, from
    the fixture), 24 independent engine ensembles x the fixture's iterations vs the reference's
    run recorded in the fixture and the oracle (pinned byte-for-byte to the reference) over 24
    seeds; tolerance 5 standard errors of the seed-to-seed scatter.  c1 = simplesin on the
    reference's light curve; c4 = pulse_vrot, a model with a prior (the swap and revert quirks act)."""
    fx = load(name)
    rows = [tuple(r) for r in fx["rows"]]
    n_par, n_beta, n_ens, iters = len(rows), fx["config"]["N_BETA"], 24, fx["config"]["MAX_ITERATIONS"]
    data = (np.loadtxt(os.path.join(GOLDEN, fx["data_file"])) if fx["data_file"]
            else np.array(fx["data"], dtype=float).reshape(-1, 2))
    cal = np.array(fx["phases"]["calibrate_rest"].split(), dtype=float).reshape(n_beta, 1 + 2 * n_par)
    n_swap = 2000 // n_beta

    def run(eng):
        eng.set_data(data)
        pt_flow.setup_chains(eng, rows)
        pt_flow.apply_calibration(eng, 0, np.tile(cal[:, 0], n_ens), np.tile(cal[:, 1:1 + n_par], (n_ens, 1)),
                                  np.tile(cal[:, 1 + n_par:], (n_ens, 1)))
        eng.reset_stats()
        eng.run(-(-iters // n_swap), n_swap)
        s = eng.get_stats()
        mean_dl = (s["sum_dl"] / s["n"]).reshape(n_ens, n_beta)
        lnz = np.array([pt_flow.evidence(cal[:, 0], m) for m in mean_dl])
        mean_p = (s["sum_params"] / s["n"][:, None]).reshape(n_ens, n_beta, n_par)[:, 0, :]
        return lnz, mean_p

    lnz_g, mp_g = run(capi.Engine(fx["model"], n_ens, n_beta, n_par=n_par, seed=101))
    lnz_c, mp_c = run(Oracle(fx["model"], n_ens, n_beta, n_par=n_par, seed=202, rng=RNG_PHILOX))
    se = np.sqrt(lnz_g.var(ddof=1) / n_ens + lnz_c.var(ddof=1) / n_ens)
    assert abs(lnz_g.mean() - lnz_c.mean()) < 5 * se, (lnz_g.mean(), lnz_c.mean(), se)
    # the reference's own single run (GSL MT19937 stream) must be a typical member
    assert abs(float(fx["evidence"]) - lnz_g.mean()) < 5 * lnz_g.std(ddof=1)
    se_p = np.sqrt(mp_g.var(axis=0, ddof=1) / n_ens + mp_c.var(axis=0, ddof=1) / n_ens)
    assert (np.abs(mp_g.mean(axis=0) - mp_c.mean(axis=0)) < 5 * se_p).all()


STAT_N_ENS = 24
STAT_TOL = 4.0   # standard errors of the difference of the two means (engine ensembles vs reference seeds)


@pytest.mark.parametrize("name", ["c1_stats", "c4_stats"])
def test_full_config_statistics_match_reference(capi, name):
    """Statistical parity at the configurations' REAL sizes (north_star: "posterior means/variances and
    the thermodynamic-integration log-evidence ... within stated statistical tolerances"):
      c1_stats  C1: simplesin on tests/testlc.dat, N_BETA 20, 20 000 iterations
      c4_stats  C4: pulse_vrot on the 2000-bin spectrum, N_BETA 20, 100 000 iterations
    The fixture (tests/golden/make_stats_golden.py) holds 24 runs of the UNMODIFIED reference (one
    thread, GSL_RNG_SEED 1..24) from ONE calibration_results: ln Z as its `analyse` prints it, and the
    mean and variance of every parameter of chain 0 over the run.  The engine runs 24 independent
    ensembles from the same calibration_results.  Stated tolerance: for ln Z, every posterior mean and
    every posterior VARIANCE, |mean over ensembles - mean over reference seeds| < 4 standard errors
    of that difference (sqrt(s_engine^2 / 24 + s_reference^2 / 24), s = scatter between runs)."""
    fx = load(name)
    rows = [tuple(r) for r in fx["rows"]]
    n_par, n_beta, iters = len(rows), fx["config"]["N_BETA"], fx["config"]["MAX_ITERATIONS"]
    n_ens = STAT_N_ENS
    data = (np.loadtxt(os.path.join(GOLDEN, fx["data_file"])) if fx["data_file"]
            else np.array(fx["data"], dtype=float).reshape(-1, 2))
    cal = np.array(fx["calibration_results"].split(), dtype=float).reshape(n_beta, 1 + 2 * n_par)
    n_swap = 2000 // n_beta
    eng = capi.Engine(fx["model"], n_ens, n_beta, n_par=n_par, seed=20261018)
    eng.set_data(data)
    pt_flow.setup_chains(eng, rows)
    pt_flow.apply_calibration(eng, 0, np.tile(cal[:, 0], n_ens), np.tile(cal[:, 1:1 + n_par], (n_ens, 1)),
                              np.tile(cal[:, 1 + n_par:], (n_ens, 1)))
    eng.reset_stats()
    left = -(-iters // n_swap)
    while left > 0:   # in calls of bounded device time
        rounds = min(left, 50)
        eng.run(rounds, n_swap)
        left -= rounds
    s = eng.get_stats()
    assert (s["n"] == -(-iters // n_swap) * n_swap).all()
    mean_dl = (s["sum_dl"] / s["n"]).reshape(n_ens, n_beta)
    lnz = np.array([pt_flow.evidence(cal[:, 0], m) for m in mean_dl])
    mean_p = (s["sum_params"] / s["n"][:, None]).reshape(n_ens, n_beta, n_par)[:, 0, :]
    var_p = (s["sum_params_sq"] / s["n"][:, None]).reshape(n_ens, n_beta, n_par)[:, 0, :] - mean_p ** 2
    runs = fx["runs"]
    assert len(runs) >= 8
    ref = dict(lnz=np.array([r["lnz"] for r in runs]), mean=np.array([r["mean"] for r in runs]),
               var=np.array([r["var"] for r in runs]))
    report = {}
    for key, ours in (("lnz", lnz), ("mean", mean_p), ("var", var_p)):
        theirs = ref[key]
        se = np.sqrt(ours.var(axis=0, ddof=1) / len(ours) + theirs.var(axis=0, ddof=1) / len(theirs))
        z = (ours.mean(axis=0) - theirs.mean(axis=0)) / se
        report[key] = np.round(np.atleast_1d(z), 2).tolist()
        assert (np.abs(z) < STAT_TOL).all(), (name, key, ours.mean(axis=0), theirs.mean(axis=0), se, z)
    print(name, "z-scores (engine - reference, in standard errors):", report,
          "ln Z engine %.4f +- %.4f reference %.4f +- %.4f" % (lnz.mean(), lnz.std(ddof=1) / np.sqrt(n_ens),
                                                                ref["lnz"].mean(), ref["lnz"].std(ddof=1) / np.sqrt(len(runs))))


def test_c3_shape_trajectory_equals_oracle(capi):
    """BASELINE config C3's own shape -- 64 ensembles x 64 rungs = 4096 chains on the 1M-row light
    curve, the tiled path's persistent likelihood kernel with its row splits -- against the oracle,
    trajectory for trajectory: 2 rounds of 3 Metropolis steps + swap.  The oracle steps a sample of
    the ensembles (a chain's stream depends on its global id only, so ensemble e of the engine =
    an oracle built with that ensemble's id offsets)."""
    bench = __import__("bench")
    data = bench.light_curve()
    n_ens, n_beta = bench.N_ENS, bench.N_BETA
    st = bench.chain_state(n_ens, n_beta, 1000)
    eng = capi.Engine("simplesin5", n_ens, n_beta, seed=1)
    eng.set_data(data)
    eng.set_bounds(bench.LO, bench.HI)
    eng.set_chains(0, eng.n_chains, **st)
    eng.run(2, 3, prob_every=1, params_chains=2)
    assert eng.last_path() == 1
    tr, out = eng.read_trace(), eng.get_chains()
    for e in (0, 37, 63):
        sl = slice(e * n_beta, (e + 1) * n_beta)
        o = Oracle("simplesin5", 1, n_beta, seed=1, rng=RNG_PHILOX, chain_id_offset=e * n_beta, ensemble_id_offset=e)
        o.set_data(data)
        o.set_bounds(bench.LO, bench.HI)
        o.set_chains(0, n_beta, **{k: v[sl] for k, v in st.items()})
        o.run(2, 3, prob_every=1, params_chains=2)
        tro, outo = o.read_trace(), o.get_chains()
        _compare_state({k: v[sl] for k, v in out.items()}, outo)
        np.testing.assert_allclose(tr["prob"][:, sl], tro["prob"], rtol=RTOL_TRAJ)
        np.testing.assert_allclose(tr["params"][:, sl, :], tro["params"], rtol=RTOL_TRAJ)
