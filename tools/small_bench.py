#!/usr/bin/env python
"""Chain-steps/s of the small configurations of BASELINE.json (C1 simplesin on the reference's
own light curve, C2 normal / 64 chains, C4 pulse_vrot) on the three kernel paths, beside the CPU oracle.
  python tools/small_bench.py            (on the GPU box) -> gpurun_out/small_bench.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pt_flow  # noqa: E402
from apemost_b200 import capi  # noqa: E402
from oracle_binding import Oracle, RNG_PHILOX  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pulse_spectrum(n, seed=4242):
    """SURVEY 8d's C4 input: n bins on [90, 110], two modes (the second a rotationally split triplet)"""
    rng = np.random.default_rng(seed)
    nu = np.linspace(90, 110, n)

    def lor(d, h):
        return h / (1 + (2 * np.pi * d * 0.5) ** 2)
    y = lor(98 - nu, 5.0) + lor(103 - nu, 3.0) + lor(103 - nu - 0.4, 3.0) + lor(103 - nu + 0.4, 3.0) + 0.05
    return np.stack([nu, y * rng.exponential(1.0, n)], axis=1)


def pulse_modes_case(n_modes, n_bins, n_beta):
    """apps/pulse.c with n_modes Lorentzians (n_par = 2 + 2 n_modes) on n_bins bins; no fixture:
    a fixed ladder and step widths that give a mid-range acceptance rate"""
    modes = [(92.0 + 2.4 * j, 5.0 - 0.5 * j) for j in range(n_modes)]
    rng = np.random.default_rng(99)
    nu = np.linspace(90, 110, n_bins)
    y = sum(h / (1 + (2 * np.pi * (f - nu) * 0.5) ** 2) for f, h in modes)
    data = np.stack([nu, y * rng.exponential(1.0, n_bins)], axis=1)
    rows = [(0.5, 0.05, 2.0, "lifetime", 0.01), (0.0, -1.0, 1.0, "offset", 0.01)]
    for j, (f, h) in enumerate(modes):
        rows += [(f, f - 1.0, f + 1.0, "f%d" % j, 0.005), (h, 0.0, 20.0, "h%d" % j, 0.05)]
    beta = pt_flow.chebyshev_ladder(n_beta, 0.05)
    steps = np.array([r[4] for r in rows])[None, :] * beta[:, None] ** -0.5
    params = np.tile(np.array([r[0] for r in rows]), (n_beta, 1))
    return "pulse", rows, data, beta, steps, params


def case(name, n_ens, n_beta=None):
    name, _, rows_override = name.partition(":")
    if name.startswith("pulse_modes"):
        return pulse_modes_case(int(name[len("pulse_modes"):]), int(rows_override), n_beta)
    fx = json.load(open(os.path.join(GOLDEN, name + ".json")))
    rows = [tuple(r) for r in fx["rows"]]
    data = (np.loadtxt(os.path.join(GOLDEN, fx["data_file"])) if fx["data_file"]
            else np.array(fx["data"], dtype=float).reshape(-1, 2))
    if rows_override:
        data = pulse_spectrum(int(rows_override))
    nb0 = fx["config"]["N_BETA"]
    cal = np.array(fx["phases"]["calibrate_rest"].split(), dtype=float).reshape(nb0, 1 + 2 * len(rows))
    n_beta = n_beta or nb0
    beta = pt_flow.chebyshev_ladder(n_beta, cal[-1, 0])
    steps = cal[0, 1:1 + len(rows)][None, :] * beta[:, None] ** -0.5
    params = np.tile(cal[0, 1 + len(rows):], (n_beta, 1))
    return fx["model"], rows, data, beta, steps, params


def time_engine(eng, rows, data, beta, steps, params, n_ens, rounds, n_swap):
    eng.set_data(data)
    pt_flow.setup_chains(eng, rows)
    pt_flow.apply_calibration(eng, 0, np.tile(beta, n_ens), np.tile(steps, (n_ens, 1)), np.tile(params, (n_ens, 1)))
    eng.run(1, n_swap)
    t0 = time.perf_counter()
    eng.run(rounds, n_swap)
    dt = time.perf_counter() - t0
    return eng.n_chains * rounds * n_swap / dt


def main():
    if os.environ.get("APM_LIB"):  # a variant build
        capi._lib = capi.load_library(os.environ["APM_LIB"])
    out = {}
    only = os.environ.get("SMALL_BENCH_ONLY", "")
    for label, name, n_ens, n_beta in [("C1 simplesin testlc.dat 1x20", "c1_phases", 1, 20),
                                       ("C1 simplesin testlc.dat 8x20", "c1_phases", 8, 20),
                                       ("C1 simplesin testlc.dat 64x20", "c1_phases", 64, 20),
                                       ("C2 normal 1x64", "c2_phases", 1, 64),
                                       ("C2 normal 64x64", "c2_phases", 64, 64),
                                       ("C4 pulse_vrot 200 rows 1x20", "c4_phases", 1, 20),
                                       ("C4 pulse_vrot 200 rows 64x20", "c4_phases", 64, 20),
                                       ("C4 pulse_vrot 2000 rows 1x20", "c4_phases:2000", 1, 20),
                                       ("C4 pulse_vrot 2000 rows 8x20", "c4_phases:2000", 8, 20),
                                       ("pulse 3 modes 2000 rows 1x20", "pulse_modes3:2000", 1, 20),
                                       ("pulse 3 modes 2000 rows 8x20", "pulse_modes3:2000", 8, 20),
                                       ("pulse 7 modes 200 rows 64x20", "pulse_modes7:200", 64, 20)]:
        if only and only not in label:
            continue
        model, rows, data, beta, steps, params = case(name, n_ens, n_beta)
        n_swap = max(1, 2000 // n_beta)
        res = {}
        for pname, path in (("tiled", 1), ("fused", 2), ("cluster", 3)):
            if path == 3 and model == "normal":
                continue  # data-free: nothing to spread over a cluster
            e = capi.Engine(model, n_ens, n_beta, n_par=len(rows), seed=1, path=path)
            res[pname] = time_engine(e, rows, data, beta, steps, params, n_ens, 20 if path == 1 else 200, n_swap)
            e.close()
        threads = os.cpu_count()
        o = Oracle(model, min(n_ens, 4), n_beta, n_par=len(rows), seed=1, rng=RNG_PHILOX, n_threads=threads)
        res["cpu_oracle"] = time_engine(o, rows, data, beta, steps, params, min(n_ens, 4), 5, n_swap)
        res["cpu_threads"] = threads
        o1 = Oracle(model, 1, n_beta, n_par=len(rows), seed=1, rng=RNG_PHILOX, n_threads=1)
        res["cpu_oracle_1thread"] = time_engine(o1, rows, data, beta, steps, params, 1, 5, n_swap)
        out[label] = res
        print(label, {k: (f"{v:.4g}" if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "small_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
