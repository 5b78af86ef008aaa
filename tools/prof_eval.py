#!/usr/bin/env python
"""One likelihood launch on config C3's shape (4096 parameter vectors x 1M rows) through
apm_gpu_eval -- the smallest command that exercises the hot kernel; used under ncu.
  python tools/prof_eval.py [path/to/lib.so] [reps] [model]
model = simplesin5 (default) or pulse_vrot (config C4's model on a 1M-bin synthetic spectrum: the
capture that shows the FP64 / MUFU pipe shares of a row term with a division and a logarithm)"""
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
model = sys.argv[3] if len(sys.argv) > 3 else "simplesin5"
rng = np.random.default_rng(0)
if model == "pulse_vrot":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import small_bench  # noqa: E402
    data = small_bench.pulse_spectrum(1_000_000)
    truth = np.array([0.5, 0.05, 0.4, 98.0, 5.0, 103.0, 3.0])
    params = truth[None, :] * (1 + rng.normal(0, 1e-3, (4096, 7)))
else:
    data = bench.light_curve()
    params = bench.TRUTH[None, :] + rng.normal(0, 1e-3, (4096, 4))
e = capi.Engine(model, 1, 1)
e.set_data(data)
for _ in range(reps):
    prob, _p = e.eval(params, np.ones(4096))
    print("loglik kernel ms:", e.last_kernel_ms()[0], "prob[0] =", prob[0])
