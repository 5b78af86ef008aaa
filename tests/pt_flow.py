"""Host-side phase logic used by the tests, written over the shared engine API
(apemost_b200.capi.EngineBase), so the same driver runs the CUDA engine and the
CPU oracle.  Restates the reference's phase drivers:

  calibrate_first   src/parallel_tempering.c:78-95
  calibrate_rest    src/parallel_tempering.c:115-207
  run               src/parallel_tempering.c:209-250, 347-419
  file formats      src/parallel_tempering_config.c:28-47,130-202, src/mcmc_dump.c:60-88

(the product's C host layer in apemost_b200/host implements the same flow for
the apps/*.exe front-ends; this Python twin exists so that parity tests can
drive both engines identically.)
"""
import os

import numpy as np

DEFAULTS = dict(burn_in_iterations=10000, desired_acceptance_rate=0.5, max_ar_deviation=0.01,
                iter_limit=100000, mul=0.85, adjust_step=0.5)


def fmt15(v):
    return "%.15e" % v


def parse_params_rows(rows):
    """rows of (start, min, max, name, step) -> arrays, with the reference's auto step
    (step < 0 -> 0.1 * (max - min), src/mcmc_parser.c:83-86)."""
    start = np.array([r[0] for r in rows], dtype=float)
    lo = np.array([r[1] for r in rows], dtype=float)
    hi = np.array([r[2] for r in rows], dtype=float)
    names = [r[3] for r in rows]
    step = np.array([r[4] if r[4] >= 0 else (r[2] - r[1]) * 0.1 for r in rows], dtype=float)
    return start, lo, hi, names, step


def setup_chains(eng, rows):
    """setup_chains(): every chain gets the params file's start values, beta = 1, and the
    mcmc_init defaults prob = prob_best = -1e10, prior = 0 (src/mcmc.c:37-78)."""
    start, lo, hi, names, step = parse_params_rows(rows)
    n = eng.n_chains
    eng.set_bounds(lo, hi)
    z = np.zeros(n, dtype=np.uint64)
    zz = np.zeros((n, eng.n_par), dtype=np.uint64)
    eng.set_chains(0, n, beta=np.ones(n), params=np.tile(start, (n, 1)), steps=np.tile(step, (n, 1)),
                   prob=np.full(n, -1e10), prior=np.zeros(n), prob_best=np.full(n, -1e10),
                   params_best=np.tile(start, (n, 1)), accept=z, reject=z, params_accepts=zz,
                   params_rejects=zz, n_iter=z, swapcount=z)
    return start, lo, hi, names, step


def calc_model_chain(eng, g):
    """calc_model(chains[g], NULL): evaluate the chain's current point and store prob/prior."""
    st = eng.get_chains(g, 1, fields=("params", "beta"))
    prob, prior = eng.eval(st["params"], st["beta"])
    eng.set_chains(g, 1, prob=prob, prior=prior)


def write_calibration_results(path, beta, steps, params):
    with open(path, "w") as f:
        for j in range(len(beta)):
            f.write(fmt15(beta[j]))
            for v in steps[j]:
                f.write("\t" + fmt15(v))
            for v in params[j]:
                f.write("\t" + fmt15(v))
            f.write("\n")


def read_calibration_results(path, n_chains, n_par):
    vals = np.array(open(path).read().split(), dtype=float).reshape(-1, 1 + 2 * n_par)[:n_chains]
    return vals[:, 0].copy(), vals[:, 1:1 + n_par].copy(), vals[:, 1 + n_par:].copy()


def apply_calibration(eng, first, beta, steps, params):
    """read_calibration_file(): beta (which zeroes swapcount), steps, params, params_best = params."""
    eng.set_chains(first, len(beta), beta=beta, steps=steps, params=params, params_best=params)


def calibrate_first(eng, rows, workdir=None, **cal):
    cfg = dict(DEFAULTS, **cal)
    setup_chains(eng, rows)
    calc_model_chain(eng, 0)
    sel = np.zeros(eng.n_chains, dtype=np.uint8)
    sel[0] = 1
    status, prog = eng.calibrate(select=sel, progress_capacity=200000, **cfg)
    st = eng.get_chains(0, 1)
    if workdir:
        write_calibration_results(os.path.join(workdir, "calibration_results"),
                                  st["beta"], st["steps"], st["params"])
    return st, status, prog


def chebyshev_ladder(n_beta, beta_0):
    """get_chain_beta with the default BETA_ALIGNMENT chebyshev_beta
    (src/parallel_tempering_beta.c:66-69,85-90)."""
    if n_beta == 1:
        return np.ones(1)
    i = n_beta - np.arange(n_beta) - 1
    return beta_0 + (1 - beta_0) / 2 * (1 - np.cos(i * np.pi / (n_beta - 1)))


def calc_beta_0(lo, hi, steps0, factors):
    r = (hi - lo) * 1.0
    r = r / steps0
    r = r / factors
    return float(np.max(r)) ** -0.5


def calibrate_rest(eng, rows, workdir, beta_0=-0.001, skip_calibrate_allchains=False, **cal):
    """One ensemble (n_ensembles == 1), exactly the reference's order of operations."""
    cfg = dict(DEFAULTS, **cal)
    n_beta, n_par = eng.n_beta, eng.n_par
    start, lo, hi, names, step = setup_chains(eng, rows)
    b, s, p = read_calibration_results(os.path.join(workdir, "calibration_results"), 1, n_par)
    apply_calibration(eng, 0, b, s, p)
    steps0, best0 = s[0], p[0]
    factors = np.ones(n_par)
    if n_beta > 1:
        i = 1
        b0 = calc_beta_0(lo, hi, steps0, factors) if beta_0 < 0 else beta_0
        beta1 = chebyshev_ladder(n_beta, b0)[i]
        steps1 = steps0 * beta1 ** -0.5
        eng.set_chains(i, 1, beta=[beta1], steps=[steps1], params=[best0])
        calc_model_chain(eng, i)
        sel = np.zeros(eng.n_chains, dtype=np.uint8)
        sel[i] = 1
        eng.calibrate(select=sel, **cfg)
        st1 = eng.get_chains(i, 1, fields=("steps", "beta"))
        factors = factors * beta1 ** -0.5
        factors = factors * steps0
        factors = factors / st1["steps"][0]
    if beta_0 < 0:
        beta_0 = calc_beta_0(lo, hi, steps0, factors)
    ladder = chebyshev_ladder(n_beta, beta_0)
    if n_beta > 1:
        sel = np.ones(eng.n_chains, dtype=np.uint8)
        sel[0] = 0
        for i in range(1, n_beta):
            steps_i = steps0 * ladder[i] ** -0.5
            steps_i = steps_i * factors
            eng.set_chains(i, 1, beta=[ladder[i]], steps=[steps_i], params=[best0])
        # calc_model for every chain 1..n-1 at its new beta
        st = eng.get_chains(1, n_beta - 1, fields=("params", "beta"))
        prob, prior = eng.eval(st["params"], st["beta"])
        eng.set_chains(1, n_beta - 1, prob=prob, prior=prior)
        eng.calibrate(select=sel, **dict(cfg, skip_calibrate=skip_calibrate_allchains))
    st = eng.get_chains(0, n_beta)
    write_calibration_results(os.path.join(workdir, "calibration_results"),
                              st["beta"], st["steps"], st["params"])
    return st, beta_0, factors


def run(eng, rows, workdir, max_iterations, n_swap=-30, write_files=True, names=None):
    """prepare_and_run_sampler + run_sampler for one ensemble, writing the reference's
    dump files when asked."""
    n_beta, n_par = eng.n_beta, eng.n_par
    start, lo, hi, pnames, step = setup_chains(eng, rows)
    b, s, p = read_calibration_results(os.path.join(workdir, "calibration_results"), n_beta, n_par)
    apply_calibration(eng, 0, b, s, p)
    if n_swap < 0:
        n_swap = 2000 // n_beta
    n_rounds = -(-max_iterations // n_swap)
    eng.run(n_rounds, n_swap, prob_every=1, params_chains=1)
    tr = eng.read_trace()
    if write_files:
        for k in range(n_beta):
            with open(os.path.join(workdir, f"prob-chain{k}.dump"), "w") as f:
                for a, d in zip(tr["prob"][:, k], tr["prob_minus_prior"][:, k]):
                    f.write("%6e\t%6e\n" % (a, d))
        for j, name in enumerate(pnames):
            with open(os.path.join(workdir, f"{name}-chain-0.prob.dump"), "w") as f:
                for v in tr["params"][:, 0, j]:
                    f.write("%.15e\n" % v)
    return tr, n_swap


def evidence(beta, mean_dl):
    """analyse_data_probability's rectangle rule (src/analyse.c:50-93): mean_dl[k] = mean of
    (prob - prior) of chain k; ln Z = sum_k mean_dl[k] / beta[k] * (beta[k] - beta[k+1]), beta[n] := 0."""
    beta = np.asarray(beta, dtype=float)
    mean_dl = np.asarray(mean_dl, dtype=float)
    prev = np.concatenate([beta[1:], [0.0]])
    return float(np.sum(mean_dl / beta * (beta - prev)))
