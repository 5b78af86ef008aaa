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
for path in [w for w in which if w <= 4]:
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
    if path == 3:   # calibration asked for as "cluster": the warp-group kernel, a few chains selected
        sel = np.zeros(e.n_chains, dtype=np.uint8)
        sel[[0, 1, 10]] = 1
        e.calibrate(select=sel, burn_in_iterations=400, iter_limit=2000, raise_on_failure=False)
        e.steps(4, 20, select=sel)
    e.set_marginals(2, n_bins=20, batch_size=5, max_batches=8)   # marginal statistics in the recording code
    e.run(2, 7)
    out, m = e.get_chains(), e.get_marginals()
    print("path", path, "->", e.last_path(), "accepts", int(out["accept"].sum()), "histogram entries",
          int(m["counts"].sum()), flush=True)
    e.close()
if 5 in which or len(sys.argv) == 1:
    # a data-free model: free_run_kernel (producers / deciders / book-keepers), adapt and marginals on
    e = capi.Engine("normal", 3, 64, seed=1, path=2)
    e.set_data(np.zeros((2, 2)))
    e.set_bounds([0.0], [60.0])
    n = e.n_chains
    e.set_chains(0, n, beta=np.tile(np.linspace(1.0, 0.5, 64), 3), params=np.full((n, 1), 20.0),
                 steps=np.full((n, 1), 4.0), params_best=np.full((n, 1), 20.0))
    e.set_adapt(True, 0.5)
    e.set_marginals(1, n_bins=20, batch_size=5, max_batches=30)
    e.run(4, 31, prob_every=1, params_chains=1)
    e.calibrate(burn_in_iterations=400, iter_limit=2000, raise_on_failure=False)
    out = e.get_chains()
    print("data-free ->", e.last_path(), "accepts", int(out["accept"].sum()), flush=True)
    e.close()
