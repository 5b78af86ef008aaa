#!/usr/bin/env python
"""Config C2 as BASELINE.json words it (normal, ONE ensemble of 64 chains) on the fused path -- the
command profiled under ncu for the data-free branch of fused_run_kernel.
  python tools/prof_c2.py [n_ens=1] [rounds=200]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import small_bench  # noqa: E402
from apemost_b200 import capi  # noqa: E402

if os.environ.get("APM_LIB"):  # a variant build (build_variants/)
    capi._lib = capi.load_library(os.environ["APM_LIB"])
n_ens = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 200
model, rows, data, beta, steps, params = small_bench.case("c2_phases", n_ens, 64)
e = capi.Engine(model, n_ens, 64, n_par=len(rows), seed=1, path=2)
for _ in range(3):
    rate = small_bench.time_engine(e, rows, data, beta, steps, params, n_ens, rounds, 31)
    print("path", e.last_path(), "chain-steps/s %.4g" % rate, "device ms", e.last_kernel_ms()[2])
