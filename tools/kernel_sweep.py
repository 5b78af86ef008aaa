#!/usr/bin/env python
"""Build variants of the likelihood kernel (compile-time knobs) and time each on config C3's
shape: 4096 parameter vectors x 1M rows through apm_gpu_eval.  Usage:
  python tools/kernel_sweep.py build      (here, no GPU needed)
  python tools/kernel_sweep.py run        (on the GPU box)"""
import ctypes as C
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "build_variants")
VARIANTS = {
    "base": [],
    "t384_c8u2": ["-DAPM_LL_THREADS=384", "-DAPM_LL_RPT=8", "-DAPM_LL_STAGES=4"],
    "t384_c8u1": ["-DAPM_LL_U=1", "-DAPM_LL_THREADS=384", "-DAPM_LL_RPT=8", "-DAPM_LL_STAGES=4"],
    "t384_c6u2": ["-DAPM_LL_C=6", "-DAPM_LL_THREADS=384", "-DAPM_LL_RPT=8", "-DAPM_LL_STAGES=4"],
    "t512_c8u1": ["-DAPM_LL_U=1", "-DAPM_LL_THREADS=512", "-DAPM_LL_RPT=8"],
    "t512_c4u2": ["-DAPM_LL_C=4", "-DAPM_LL_THREADS=512", "-DAPM_LL_RPT=8"],
    "t256_s4_rpt12": ["-DAPM_LL_RPT=12", "-DAPM_LL_STAGES=4"],
    "unroll2": ["-DAPM_LL_UNROLL=2"],
    "unroll8": ["-DAPM_LL_UNROLL=8"],
}
# run-time knobs tried on the base build
ENVS = [{}, {"APM_SPLITS": "13"}, {"APM_SPLITS": "26"}]


def build():
    import __graft_entry__ as g
    os.makedirs(OUT, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor

    def one(item):
        name, flags = item
        out = os.path.join(OUT, f"libapm_{name}.so")
        g.build_cuda(force=True, extra_flags=flags, out=out)
        print("built", name, flush=True)
    with ThreadPoolExecutor(max_workers=int(os.environ.get("JOBS", "4"))) as ex:
        list(ex.map(one, VARIANTS.items()))


def run():
    sys.path.insert(0, ROOT)
    import bench
    from apemost_b200 import capi
    data = bench.light_curve()
    rng = np.random.default_rng(0)
    params = bench.TRUTH[None, :] + rng.normal(0, 1e-3, (4096, 4))
    beta = np.ones(4096)
    results = {}
    envs = ENVS
    for name in VARIANTS:
        path = os.path.join(OUT, f"libapm_{name}.so")
        if not os.path.exists(path):
            continue
        for env in (envs if name == "base" else [{}]):
            for k in ("APM_SPLITS",):
                os.environ.pop(k, None)
            os.environ.update(env)
            lib = capi.load_library(path)
            capi._lib = lib
            e = capi.Engine("simplesin5", 1, 1)
            e.set_data(data)
            best = 1e9
            ref = None
            for rep in range(4):
                prob, _ = e.eval(params, beta)
                ms, n, tot = e.last_kernel_ms()
                best = min(best, ms)
            key = name + ("" if not env else "+" + ",".join(f"{k}={v}" for k, v in env.items()))
            results[key] = {"loglik_ms": best, "rowevals_per_s": 4096e6 / (best * 1e-3), "prob0": float(prob[0])}
            print(key, results[key], flush=True)
            e.close()
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
