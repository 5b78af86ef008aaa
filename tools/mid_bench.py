#!/usr/bin/env python
"""Mid-size tables (too large for shared memory, far too small to fill the GPU for long): chain-steps/s
of simplesin5 on 20-rung ladders for a range of table sizes and ensemble counts, with the per-step
device time.  python tools/mid_bench.py   (on the GPU box)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from apemost_b200 import capi  # noqa: E402

for n_rows in (20_000, 100_000, 400_000, 1_000_000):
    data = bench.light_curve(n_rows)
    for n_ens in (1, 8, 24):
        n_beta = 20
        st = bench.chain_state(n_ens, n_beta, 5)
        st["steps"] = st["steps"] * (1e6 / n_rows) ** 0.5
        for path, name in ((1, "tiled"), (4, "grid")):
            e = capi.Engine("simplesin5", n_ens, n_beta, seed=1, path=path)
            e.set_data(data)
            e.set_bounds(bench.LO, bench.HI)
            e.set_chains(0, e.n_chains, **st)
            try:
                e.run(2, 100)
            except Exception as ex:
                print(f"rows {n_rows:7d}  ensembles {n_ens:3d}  {name}: {str(ex)[:80]}")
                continue
            rounds = 10
            t0 = time.perf_counter()
            e.run(rounds, 100)
            dt = time.perf_counter() - t0
            ll_ms, n_ll, total_ms = e.last_kernel_ms()
            print(f"rows {n_rows:7d}  ensembles {n_ens:3d}  {name:5s}  {e.n_chains * rounds * 100 / dt:12.4g} chain-steps/s  "
                  f"{total_ms / (rounds * 100) * 1e3:8.2f} us/step", flush=True)
            e.close()
