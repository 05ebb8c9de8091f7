#!/usr/bin/env python
"""Micro-benchmark of the covariance + CG hot path at a fixed BFGS rank: builds a DCT-basis state with `pairs`
space updates, then times `iters` CG iterations (no early exit) and the covariance updates with CUDA events.
Used for the ncu captures under profiles/ and for the per-iteration roofline numbers in DESIGN.md.

    python tools/cg_microbench.py --batch 1 --pairs 16 --operator gaussian_blur --iters 200
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from free_hunch_b200 import engine as E  # noqa: E402

D_VEC = 3 * 256 * 256 * 4


def build(args):
    assets = np.load(os.path.join(ROOT, "tests", "golden", "assets.npz"))
    var = assets["dct_variance"]
    B = args.batch
    ctx = E.Context("cuda", max_batch=B)
    st = E.CovState(ctx, B, E.BASIS_DCT)
    s0 = 80.2878
    st.init(torch.from_numpy(var), s0 ** 2)
    g = torch.Generator(device="cuda").manual_seed(0)
    vt = torch.from_numpy(var).cuda()
    shape = (B, 3, 256, 256)

    def den(x, s):
        return ctx.dct2(ctx.dct2(x) * (vt / (vt + s * s)), inverse=True) + 0.02 * torch.tanh(x)
    sig = np.geomspace(9.5, 1.2, max(args.pairs, 1))
    x = torch.randn(shape, device="cuda", generator=g) * s0
    s_prev = s0
    t_time = t_space = 0.0
    for i in range(args.pairs):
        s = float(sig[i])
        mu_prev = den(x, s_prev)
        score = (mu_prev - x) / s_prev ** 2
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        mean, _ = st.time_update(x, score, s_prev, s)
        e[1].record()
        xn = x + (s - s_prev) * (x - mu_prev) / s_prev + 0.02 * s * torch.randn(shape, device="cuda", generator=g)
        mun = den(xn, s)
        e[2].record()
        # secant pair from the denoiser itself at the new sigma (dx.de > 0: a positive-definite state; with the time
        # update's own prediction `mean` the pair carries the model error and the synthetic state can go indefinite,
        # which ends CG through the p.Ap <= 1e-16 break after a handful of iterations)
        st.space_update(den(x, s), mun, s, x, xn)
        e3 = torch.cuda.Event(enable_timing=True)
        e3.record()
        torch.cuda.synchronize()
        t_time, t_space = e[0].elapsed_time(e[1]), e[2].elapsed_time(e3)
        x, s_prev = xn, s
    if args.operator == "inpainting":
        mask = (torch.rand(B, 256, 256, device="cuda", generator=g) > 0.7).float()
        op = E.Operator.inpaint(ctx, mask)
    elif args.operator == "super_resolution":
        op = E.Operator.conv(ctx, assets["bicubic_x4"], 4)
    else:
        op = E.Operator.conv(ctx, assets["gaussian_ks61_std3" if args.operator == "gaussian_blur" else "motion_ks61_std05"], 1)
    y = torch.randn(op.y_shape(B), device="cuda", generator=g)
    mu = torch.randn(shape, device="cuda", generator=g) * 0.3
    return ctx, st, op, y, mu, t_time, t_space


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--pairs", type=int, default=16)
    ap.add_argument("--operator", default="gaussian_blur")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--profile-kernels", action="store_true")
    ap.add_argument("--f64", action="store_true", help="float64 y / mu / mat (fh_cg_solve_f64)")  # float64_io
    ap.add_argument("--mode", default="auto", choices=["auto", "streaming", "resident", "resident_f32"])
    args = ap.parse_args()
    ctx, st, op, y, mu, t_time, t_space = build(args)
    if args.f64:
        y, mu = y.double(), mu.double()
    B, k = args.batch, st.rank // 2
    ctx.set_cg_mode({"auto": E.CG_AUTO, "streaming": E.CG_STREAMING, "resident": E.CG_RESIDENT, "resident_f32": E.CG_RESIDENT_F32}[args.mode])
    for _ in range(2):
        E.cg_solve(st, op, y, mu, 0.01, 1.0, maxiter=args.iters, fixed_iters=args.iters)
    torch.cuda.synchronize()
    E.launch_count(reset=True)
    if args.profile_kernels:
        E.profile_kernels(True)
    ms, wall = [], []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        mat, info = E.cg_solve(st, op, y, mu, 0.01, 1.0, maxiter=args.iters, fixed_iters=args.iters)
        e1.record()
        e1.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(e0.elapsed_time(e1))
    launches = E.launch_count()
    done = max(1, max(info["niter"]))                 # the pAp <= 1e-16 break may end the loop early
    us_it = float(np.median(ms)) / done * 1e3
    nvec = {"gaussian_blur": 7 + 2 * k, "motion_blur": 7 + 2 * k, "inpainting": 8 + 2 * k, "super_resolution": 1 + 2 * k}[args.operator]
    alg = nvec * D_VEC * B
    out = {"operator": args.operator, "batch": B, "pairs": k, "iters": args.iters, "us_per_iteration": us_it,
           "us_per_image_iteration": us_it / B, "wall_us_per_iteration": float(np.median(wall)) / done * 1e3,
           "launches_per_iteration": launches / args.reps / done,
           "algorithmic_MB_per_iteration": alg / 1e6, "achieved_GBps": alg / (us_it * 1e-6) / 1e9,
           "time_update_ms": t_time, "space_update_ms": t_space, "niter": info["niter"][0],
           "cg_path": {E.CG_STREAMING: "streaming", E.CG_RESIDENT: "resident (float64)", E.CG_RESIDENT_F32: "resident (float32)"}.get(ctx.cg_last_mode)}
    if args.profile_kernels:
        E.profile_kernels(False)
        out["kernels"] = {n: [c, round(t / c * 1e3, 2)] for n, (c, t) in E.kernel_profile().items()}
    if os.environ.get("FH_TRACE"):
        import ctypes
        from free_hunch_b200 import _lib
        buf = (ctypes.c_longlong * 128)()
        _lib.check(_lib.lib().fh_trace_read(ctx._h, buf))
        tr = np.array(list(buf)).reshape(8, 16)
        out["trace_us"] = {str(k): [round((int(v) - int(tr[k][0])) / 1965.0, 2) if v else None for v in tr[k]] for k in range(8) if tr[k][0]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
