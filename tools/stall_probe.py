#!/usr/bin/env python
"""Looks for rare long stalls: many short CG solves, each timed with CUDA events and host wall clock; prints the
distribution and the outliers.  python tools/stall_probe.py --mode resident|streaming --solves 1500"""
import argparse, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from free_hunch_b200 import engine as E  # noqa: E402
import cg_microbench as mb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="resident")
    ap.add_argument("--solves", type=int, default=1500)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--asyncs", action="store_true")
    a = ap.parse_args()
    args = argparse.Namespace(batch=1, pairs=a.pairs, operator="gaussian_blur")
    ctx, st, op, y, mu, _, _ = mb.build(args)
    ctx.set_cg_mode({"resident": E.CG_RESIDENT, "streaming": E.CG_STREAMING}[a.mode])
    dev_ms, wall_ms = [], []
    evs = []
    t_all = time.perf_counter()
    for i in range(a.solves):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if a.asyncs:
            E.cg_solve_async(st, op, y, mu, 0.01, 1.0, maxiter=a.iters, fixed_iters=a.iters)
        else:
            E.cg_solve(st, op, y, mu, 0.01, 1.0, maxiter=a.iters, fixed_iters=a.iters)
        e1.record()
        evs.append((e0, e1))
        wall_ms.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    total = (time.perf_counter() - t_all) * 1e3
    dev_ms = np.array([a_.elapsed_time(b_) for a_, b_ in evs])
    wall_ms = np.array(wall_ms)
    med = float(np.median(dev_ms))
    out = [(i, round(float(v), 2)) for i, v in enumerate(dev_ms) if v > 3 * med]
    print(f"mode {a.mode} async {a.asyncs}: {a.solves} solves x {a.iters} it, total {total:.0f} ms, device median {med:.3f} ms, p99 {np.percentile(dev_ms, 99):.3f}, "
          f"max {dev_ms.max():.2f}; wall median {np.median(wall_ms):.3f} max {wall_ms.max():.2f}; outliers(>3x median): {out[:20]}")


if __name__ == "__main__":
    main()
