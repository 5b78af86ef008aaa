import os, sys, time
import numpy as np
sys.path.insert(0,'/root/repo')
import bench
from apemost_b200 import capi
for n_rows in (20000, 100000, 400000):
    data = bench.light_curve(n_rows)
    for path, name in ((1,'tiled'),(4,'grid')):
        n_ens, n_beta = 1, 20
        st = bench.chain_state(n_ens, n_beta, 5)
        st["steps"] = st["steps"] * (1e6/n_rows)**0.5
        e = capi.Engine("simplesin5", n_ens, n_beta, seed=1, path=path)
        e.set_data(data); e.set_bounds(bench.LO, bench.HI); e.set_chains(0, e.n_chains, **st)
        prob, prior = e.eval(st["params"], st["beta"]); e.set_chains(0, e.n_chains, prob=prob, prior=prior)
        t0=time.perf_counter(); status,_ = e.calibrate(burn_in_iterations=2000, raise_on_failure=False); dt=time.perf_counter()-t0
        out=e.get_chains()
        print(f"rows {n_rows:7d} {name:5s} path {e.last_path()} calibrate 20 chains: {dt:.3f} s, status ok {int((status==0).sum())}, rng steps {int(out['rng_counter'].sum())}", flush=True)
        e.close()
