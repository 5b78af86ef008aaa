#!/usr/bin/env python
"""A small run on every kernel path (run, adapt, calibration, steps) -- a quick smoke of all four,
and the command to put under compute-sanitizer where that tool is available:
  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from apemost_b200 import capi  # noqa: E402

which = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
for path in which:
    n_rows = 1522 if path in (2, 3) else 30000
    data = bench.light_curve(n_rows)
    n_ens, n_beta = 2, 10
    st = bench.chain_state(n_ens, n_beta, 5)
    st["steps"] = st["steps"] * (1e6 / n_rows) ** 0.5
    e = capi.Engine("simplesin5", n_ens, n_beta, seed=1, path=path)
    e.set_data(data)
    e.set_bounds(bench.LO, bench.HI)
    e.set_chains(0, e.n_chains, **st)
    e.set_adapt(True, 0.5)
    e.run(3, 7, prob_every=1, params_chains=1)
    if path in (1, 2):
        e.calibrate(burn_in_iterations=400, iter_limit=2000, raise_on_failure=False)
        e.steps(1, 20)
    out = e.get_chains()
    print("path", path, "->", e.last_path(), "accepts", int(out["accept"].sum()), flush=True)
    e.close()
