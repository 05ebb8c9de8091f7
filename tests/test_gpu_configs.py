"""FULL BASELINE configurations against goldens produced by the REAL reference (tests/golden/make_golden.py::golden_config):
the whole Free Hunch mechanism inside the complete EDM loop -- 30-step Heun (59 mechanism calls) for gaussian_blur,
random inpainting and x4 super-resolution, 100-step Euler for motion_blur -- at the reference's default max_rtol = 1 (the
regime bench.py times) and at a converged max_rtol = 1e-6.  Compared call by call: CG iteration counts, every mechanism
output, the final sample.  The stand-in denoiser is the same float64 dense-DCT model the golden run used, so the runs
differ only in the hot path under test.

Tolerances (north-star: 1e-4 relative L2 for posterior mean, CG solution and final sample):
  * converged goldens: final sample AND every intermediate mechanism output within 1e-4, total CG iterations within 2 %
    (measured with the float64 covariance updates of fh_k_upd64: final 3.8e-8, worst call 1.4e-6, 32 835 vs 32 860
    iterations; with the float32 multi-launch updates before them: 3.4e-5 / 2.2e-4 inside the space-update window,
    where mu + Sigma mat cancels to ~1 % of its terms);
  * default-rtol goldens: for Gaussian blur the FIRST call (sigma = 80, the condition-number-1e6 system that float32 CG
    could not follow: 32 instead of 27 iterations, rel-L2 0.80 in round 1) must stop at the reference's iteration and agree
    to 1e-3.  Everything later -- and the first call of the other operators -- stops on `|r| <= rtol |b|` far from
    convergence, where the reference ITSELF is not reproducible: fed a denoiser mean that differs by 1.3e-15 (scipy DCT
    instead of dense DCT matrices) the reference's own inpainting solve stops after 18 instead of 23 iterations (measured
    in the build container with the real reference).  There the bar is the iteration total within 2 % and the final
    sample within that float64-vs-float64 sensitivity (5e-2).
"""
import os

import numpy as np
import pytest
import torch

import helpers as H
from helpers import rel_l2, make_inputs, sub, record

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not H.HAS_CUDA, reason="needs a CUDA device")]

from oracle import fh_oracle as fo  # noqa: E402

SHAPE = (1, 3, 256, 256)
GOLD = os.path.join(H.ROOT, "tests", "golden")
CONFIGS = {   # name -> (operator, solver)
    "c1_gauss_heun30": ("gaussian_blur", "heun"),
    "c2_inpaint_heun30": ("inpainting", "heun"),
    "c3_sr_heun30": ("super_resolution", "heun"),
    "c4_motion_euler100": ("motion_blur", "euler"),
    "c1_gauss_heun30_tight": ("gaussian_blur", "heun"),
    "c2_inpaint_heun30_tight": ("inpainting", "heun"),
    "c3_sr_heun30_tight": ("super_resolution", "heun"),
    "c4_motion_euler100_tight": ("motion_blur", "euler"),
}


def synth_image(var, seed):
    z = make_inputs(seed, var.shape) * np.sqrt(var.astype(np.float64))
    return np.clip(fo.idct2(z), -1, 1)[None]


def tt(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def build_problem(name, g, assets, dev):
    """The inputs of golden_config, rebuilt without the reference: (operator, y)."""
    from free_hunch_b200.measurements import get_operator
    operator, _ = CONFIGS[name]
    var = assets["dct_variance"]
    seed = int(g["seed"])
    x_clean = synth_image(var, seed)
    kw = dict(in_shape=SHAPE, sigma_s=0.1, device=dev)
    if operator == "inpainting":
        mask = np.unpackbits(g["mask_bits"]).reshape(256, 256).astype(np.float32)
        op = get_operator("inpainting", mask=tt(mask)[None, None].repeat(1, 3, 1, 1), **kw)
        y = (x_clean + 0.1 * make_inputs(seed + 4, SHAPE)) * mask
    elif operator == "super_resolution":
        op = get_operator("super_resolution", kernel=assets["bicubic_x4"], scale_factor=4, **kw)
        y = g["y"].astype(np.float64)
    else:
        key = "gaussian_ks61_std3" if operator == "gaussian_blur" else "motion_ks61_std05"
        op = get_operator(operator, kernel=assets[key], scale_factor=1, **kw)
        oo = (fo.GaussianBlurOracle if operator == "gaussian_blur" else fo.MotionBlurOracle)(assets[key], 0.1)
        # the measurement was synthesised with the reference's complex64 OTF
        fb = g["FB"] if "FB" in g.files else np.load(os.path.join(GOLD, "mechanism.npz"))["FB"]
        oo.FB = fb[None, None]
        y = oo.A(x_clean) + 0.1 * make_inputs(seed + 4, SHAPE)
    return op, y, make_inputs(seed - 3, SHAPE)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_config_vs_reference_golden(name, assets):
    path = os.path.join(GOLD, f"cfg_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    from free_hunch_b200.conditioning_mechanisms import choose_conditioning_mechanism
    from free_hunch_b200.sampler import AnalyticDCTGaussianDenoiser64, conditional_sampler
    g = np.load(path)
    dev = torch.device("cuda")
    var = assets["dct_variance"]
    operator, solver = CONFIGS[name]
    op, y, noise = build_problem(name, g, assets, dev)
    t_steps = g["t_steps"]
    max_rtol = float(g["max_rtol"])
    mech = choose_conditioning_mechanism("free_hunch")(
        1.0, op, False, init_denoiser_variance=1, init_noise_variance=torch.tensor(t_steps[0] ** 2, dtype=torch.float64),
        data_dim=3 * 256 * 256, max_vector_count=100000, dct_variance=var, image_base_covariance="dct_diagonal",
        denoiser_mean_error_threshold=0.2, use_analytical_score_time_update=True, project_to_diagonal=False,
        space_step_update_threshold=10.0, space_step_update_lower_threshold=1.0, max_rtol=max_rtol, do_space_updates=True,
        use_analytic_var_at_end=False, solver_type="customcuda", use_rtol_func=True, device=dev)
    net = AnalyticDCTGaussianDenoiser64(var, dev)
    outs = []

    def spy(x_t, mdl, yy, s):
        o = mech(x_t, mdl, yy, s)
        outs.append(sub(o))
        return o
    x = conditional_sampler(spy, net, tt(y).to(dev), tt(noise).to(dev), t_steps, solver=solver)
    ref_n = np.array(g["niters"])
    got_n = np.array([r["niter"][0] for r in mech.log])
    errs = np.array([rel_l2(o, g[f"call{i}_out"]) for i, o in enumerate(outs)])
    ef = rel_l2(sub(x), g["final"])
    ncalls = len(ref_n)
    first_diff = int(np.argmax(got_n != ref_n)) if np.any(got_n != ref_n) else -1
    record("full_config", cfg=name, max_rtol=max_rtol, calls=ncalls, final=ef, first_call=float(errs[0]), worst_call=float(errs.max()),
           worst_call_index=int(errs.argmax()), cg_total=int(got_n.sum()), ref_cg_total=int(ref_n.sum()),
           calls_with_same_niter=int((got_n == ref_n).sum()), first_call_with_other_niter=first_diff,
           k=mech.covariance_model.vector_count, ref_k=int(g["k_final"]),
           niter=got_n.tolist(), ref_niter=ref_n.tolist(), errs=[float(f"{e:.3g}") for e in errs])
    assert len(got_n) == ncalls
    assert mech.covariance_model.vector_count == int(g["k_final"])
    assert abs(int(got_n.sum()) - int(ref_n.sum())) <= 0.02 * ref_n.sum(), (got_n.sum(), ref_n.sum())
    if max_rtol < 1e-3:
        assert ef < 1e-4, ef
        assert errs.max() < 1e-4, (errs.argmax(), errs.max())
    else:
        if operator == "gaussian_blur":
            assert got_n[0] == ref_n[0] and errs[0] < 1e-3, (got_n[0], ref_n[0], errs[0])
        assert ef < 5e-2, ef
