#!/usr/bin/env python
"""Runs the CPU oracle over the whole default bench workload once (image seed 1000, gaussian_blur, 30-step Heun) and
stores its per-call CG iteration counts in tests/golden/bench_oracle_iters.json.  bench.py uses the total to scale its
BOUNDED CPU sample (a prefix of the calls) to images/s by CG-iteration share instead of by call count."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class A:
    operator = "gaussian_blur"; num_steps = 30; solver = "heun"


if __name__ == "__main__":
    assets = np.load(os.path.join(ROOT, "tests", "golden", "assets.npz"))
    out = {}
    for op, steps, solver in [("gaussian_blur", 30, "heun")]:
        a = A(); a.operator, a.num_steps, a.solver = op, steps, solver
        t0 = time.time()
        calls, complete, secs = bench.cpu_reference_sample(assets, a, 1e9, os.cpu_count())
        assert complete
        out[f"{op}/{steps}/{solver}"] = {"niters": [c["niter"] for c in calls], "seconds_here": secs, "cores_here": os.cpu_count()}
        print(op, sum(c["niter"] for c in calls), secs)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bench_oracle_iters.json"), "w"))
