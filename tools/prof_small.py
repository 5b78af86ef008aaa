#!/usr/bin/env python
"""One small-table run (config C1's shape: simplesin on the reference's light curve, 20-rung
ladder) on a chosen kernel path -- the command profiled under ncu for the fused / cluster kernels.
  python tools/prof_small.py [path=3] [n_ens=1] [rounds=20] [n_swap=100]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import small_bench  # noqa: E402
from apemost_b200 import capi  # noqa: E402

if os.environ.get("APM_LIB"):  # a variant build, e.g. build_variants/libapm_timing.so (-DAPM_CLUSTER_TIMING)
    capi._lib = capi.load_library(os.environ["APM_LIB"])
path = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_ens = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 20
n_swap = int(sys.argv[4]) if len(sys.argv) > 4 else 100
model, rows, data, beta, steps, params = small_bench.case("c1_phases", n_ens, 20)
e = capi.Engine(model, n_ens, 20, n_par=len(rows), seed=1, path=path)
for _ in range(3):
    rate = small_bench.time_engine(e, rows, data, beta, steps, params, n_ens, rounds, n_swap)
    print("path", e.last_path(), "chain-steps/s %.4g" % rate, "device ms", e.last_kernel_ms()[2])
