#!/usr/bin/env python
"""One likelihood launch on config C3's shape (4096 parameter vectors x 1M rows) through
apm_gpu_eval -- the smallest command that exercises the hot kernel; used under ncu.
  python tools/prof_eval.py [path/to/lib.so] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from apemost_b200 import capi  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] != "-":
    capi._lib = capi.load_library(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
data = bench.light_curve()
rng = np.random.default_rng(0)
params = bench.TRUTH[None, :] + rng.normal(0, 1e-3, (4096, 4))
e = capi.Engine("simplesin5", 1, 1)
e.set_data(data)
for _ in range(reps):
    prob, _p = e.eval(params, np.ones(4096))
    print("loglik kernel ms:", e.last_kernel_ms()[0], "prob[0] =", prob[0])
