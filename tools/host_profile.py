#!/usr/bin/env python
"""Where does the host time of one guided sampling go?  cProfile of one bench.py-style step (after warm-up)."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    opname = sys.argv[1] if len(sys.argv) > 1 else "gaussian_blur"
    from free_hunch_b200.conditioning_mechanisms import choose_conditioning_mechanism
    from free_hunch_b200.measurements import get_operator
    from free_hunch_b200.sampler import AnalyticDCTGaussianDenoiser, conditional_sampler, edm_sigma_steps
    assets = np.load(os.path.join(ROOT, "tests", "golden", "assets.npz"))
    var = assets["dct_variance"]
    dev = torch.device("cuda", 0)
    t_steps = edm_sigma_steps(30)
    x_clean, noise, ynoise = bench.synth_inputs(assets, opname, 1, 1000)
    if opname == "inpainting":
        mask = bench.random_mask(1000)
        op = get_operator("inpainting", mask=torch.from_numpy(mask)[None, None].repeat(1, 3, 1, 1), in_shape=(1, 3, 256, 256), sigma_s=0.1, device=dev)
        y = ((x_clean + 0.1 * ynoise) * mask[None, None]).astype(np.float32)
    else:
        op = get_operator(opname, kernel=assets["gaussian_ks61_std3"], scale_factor=1, in_shape=(1, 3, 256, 256), sigma_s=0.1, device=dev)
        y = (op.A(torch.from_numpy(x_clean).float().to(dev)).double().cpu().numpy() + 0.1 * ynoise).astype(np.float32)
    y_dev, n_dev = torch.from_numpy(y).to(dev), torch.from_numpy(noise.astype(np.float32)).to(dev)

    def run():
        mech = choose_conditioning_mechanism("free_hunch")(
            1.0, op, False, init_denoiser_variance=1, init_noise_variance=torch.tensor(float(t_steps[0]) ** 2, dtype=torch.float64),
            data_dim=3 * 256 * 256, max_vector_count=100000, dct_variance=var, image_base_covariance="dct_diagonal",
            denoiser_mean_error_threshold=0.2, use_analytical_score_time_update=True, project_to_diagonal=False,
            space_step_update_threshold=10.0, space_step_update_lower_threshold=1.0, max_rtol=1.0, do_space_updates=True,
            use_analytic_var_at_end=False, solver_type="customcuda", use_rtol_func=True, device=dev, batch=1, defer_host_reads=True)
        net = AnalyticDCTGaussianDenoiser(var, mech.covariance_model.ctx)
        fin = mech.finalize
        def traced_fin():
            stamp["enqueued"] = time.perf_counter()        # the host has queued the whole image; what follows is waiting
            return fin()
        mech.finalize = traced_fin
        x = conditional_sampler(mech, net, y_dev, n_dev, t_steps, solver="heun")
        torch.cuda.synchronize()
        return x
    stamp = {}
    for _ in range(3):
        run()
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run()
        t1 = time.perf_counter()
        print(f"{opname}: wall per image {(t1 - t0) * 1e3:.1f} ms, host finished queueing after {(stamp['enqueued'] - t0) * 1e3:.1f} ms")
    pr = cProfile.Profile()
    pr.enable()
    run()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue())
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
    print(s.getvalue())


if __name__ == "__main__":
    main()
