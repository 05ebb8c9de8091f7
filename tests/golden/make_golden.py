"""Generate golden vectors by running the REAL reference (``/root/reference``) on seeded inputs.

Run once, in the build container (the reference is not available on the GPU box):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` replays the same inputs through
``oracle/fh_oracle.py`` and compares.  The import recipe is SURVEY.md Appendix B: stub ``torch_dct``
with scipy's orthonormal DCT, stub absent plotting / wavelet modules, inject the missing
``CovarianceHessianBFGSDCTPCA`` symbol, make ``.cuda()`` the identity (no GPU here), chdir into the
reference root (it opens data files by relative path).  Nothing here is product code.
"""
import os
import sys
import tempfile
import types

import numpy as np
import scipy.fft as sfft
import scipy.io
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
SUB = 97  # stride of the subsample kept for full-size (196608-element) outputs


def import_reference():
    sys.path.insert(0, REF)
    os.chdir(REF)
    td = types.ModuleType("torch_dct")
    td.dct_2d = lambda x, norm=None: torch.from_numpy(
        sfft.dctn(x.detach().cpu().numpy(), type=2, norm=norm, axes=(-2, -1))).to(x.dtype)
    td.idct_2d = lambda x, norm=None: torch.from_numpy(
        sfft.idctn(x.detach().cpu().numpy(), type=2, norm=norm, axes=(-2, -1))).to(x.dtype)
    sys.modules["torch_dct"] = td
    for n in ("pywt", "hdf5storage", "matplotlib", "matplotlib.pyplot"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["hdf5storage"].loadmat = scipy.io.loadmat
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    import conditioning_utils.online_update_bfgs as ob
    ob.CovarianceHessianBFGSDCTPCA = None
    import conditioning_utils.conditioning_mechanisms as cm
    import conditioning_utils.cg as rcg
    from measurement_utils import measurements as ms
    import measurement_utils.utils_sisr as sisr
    return ob, cm, rcg, ms, sisr


def sub(a):
    return np.asarray(a, np.float64).reshape(-1)[::SUB].copy()


def dense4(bfgs):
    return [m.real.numpy().astype(np.float64) for m in bfgs.get_dense_matrices()]


# ------------------------------------------------------------------------------------------
def golden_unit(ob):
    """The reference's four unit tests (conditioning_utils/online_update_tests.py) with the F3
    adaptation: leading batch dim on the class inputs, U/V cast to complex128."""
    g = {}
    c128 = torch.complex128

    # --- test_covariance_hessian_time_update (:3-69), seed added for reproducibility
    torch.manual_seed(1234)
    d, t, tn = 5, 20.0, 18.0
    x = torch.randn(d, dtype=torch.float64)
    score = -x / t ** 2
    bf = ob.CovarianceHessianBFGS(init_denoiser_variance=1, init_noise_variance=t ** 2, data_dim=d)
    mean, sc = bf.update_time_step(x[None], t, tn, score[None])
    g["t1_x"], g["t1_score"] = x.numpy(), score.numpy()
    g["t1_mean"], g["t1_newscore"] = mean.numpy()[0], sc.numpy()[0]
    for n, m in zip("cov inv_cov hes inv_hes".split(), dense4(bf)):
        g[f"t1_{n}"] = m
    den = torch.eye(d, dtype=torch.float64)
    hes = (den / t ** 2 - torch.eye(d)) / t ** 2
    ref = ob.update_covariance(x, den, den.clone(), hes, torch.linalg.inv(hes), score, x + t ** 2 * score, lambda s: s, t, tn)
    for n, m in zip("cov inv_cov hes inv_hes newscore mean".split(), ref):
        g[f"t1_dense_{n}"] = m.numpy()

    # --- test_covariance_hessian_time_update_with_u_and_v (:71-187)
    torch.manual_seed(42)
    for pairs in [1, 2, 4, 8]:
        d, t, tn = 15, 80.0, 79.0
        x = torch.randn(d, dtype=torch.float32)
        U, V = [], []
        for _ in range(pairs):
            u = torch.randn(d); v = torch.randn(d)
            u = u / torch.norm(u); v = v / torch.norm(v)
            v = v - torch.dot(u, v) * u
            v = v / torch.norm(v) * np.sqrt(0.5)
            U.append(u[:, None]); V.append(v[:, None])
        U = torch.cat(U, 1); V = torch.cat(V, 1)
        score = -x / t ** 2
        bf = ob.CovarianceHessianBFGS(init_denoiser_variance=1, init_noise_variance=t ** 2, data_dim=d)
        bf.vectors_denoiser_cov_u = U.to(c128)
        bf.vectors_denoiser_cov_v = V.to(c128)
        bf.set_others_corresponding_to_current_denoiser_cov(t)
        p = f"t2_{pairs}_"
        g[p + "x"], g[p + "U"], g[p + "V"] = x.numpy(), U.numpy(), V.numpy()
        for n, m in zip("cov inv_cov hes inv_hes".split(), dense4(bf)):
            g[p + "pre_" + n] = m
        mean, sc = bf.update_time_step(x[None], t, tn, score[None])
        g[p + "mean"], g[p + "newscore"] = mean.numpy()[0], sc.numpy()[0]
        for n, m in zip("cov inv_cov hes inv_hes".split(), dense4(bf)):
            g[p + n] = m

    # --- test_bfgs_update (:189-272): 10 BFGS steps, d=15
    def score_fn(x, t):
        return 0.7 * (-x / t ** 2) + 0.3 * (-0.5 * (x - torch.ones_like(x)) / t ** 2)
    torch.manual_seed(0); np.random.seed(0)
    d, t = 15, 50.0
    x = torch.randn(d, dtype=torch.float32)
    bf = ob.CovarianceHessianBFGS(init_denoiser_variance=1, init_noise_variance=t ** 2, data_dim=d)
    den = torch.eye(d); inv = torch.eye(d)
    g["t3_x0"] = x.numpy()
    dxs = []
    for step in range(10):
        dx = torch.randn(d) * 0.1
        dxs.append(dx.numpy())
        xn = x + dx
        m0 = x + t ** 2 * score_fn(x, t)
        m1 = xn + t ** 2 * score_fn(xn, t)
        den, inv, hes, ihes = ob.update_bfgs(den, inv, m0[None], m1[None], lambda s: s, t, x[None], dx[None])
        bf.update_space_step(m0[None], m1[None], t, x[None], xn[None])
        for n, m in zip("cov inv_cov hes inv_hes".split(), dense4(bf)):
            g[f"t3_{step}_{n}"] = m
        g[f"t3_{step}_dense_cov"] = den.numpy()[0] if den.dim() == 3 else den.numpy()
        g[f"t3_{step}_dense_inv_cov"] = inv.numpy()[0] if inv.dim() == 3 else inv.numpy()
        x = xn
    g["t3_dxs"] = np.stack(dxs)

    # --- test_time_and_space_updates (:274-374): 25 alternating steps, d=5
    torch.manual_seed(0); np.random.seed(0)
    d = 5
    x = torch.randn(d, dtype=torch.float32).to(torch.float64)
    ts = [50.0, 48.0, 46.0, 44.0, 42.0, 40.0, 38.0, 36.0, 34.0, 32.0, 30.0, 28.0, 26.0, 24.0, 22.0, 20.0,
          18.0, 16.0, 14.0, 12.0, 10.0, 8.0, 6.0, 4.0, 2.0, 0.1]
    bf = ob.CovarianceHessianBFGS(init_denoiser_variance=1, init_noise_variance=ts[0] ** 2, data_dim=d, dtype=c128)
    g["t4_x0"], g["t4_ts"] = x.numpy(), np.array(ts)
    dxs = []
    for i in range(len(ts) - 1):
        t, tn = ts[i], ts[i + 1]
        dx = (torch.randn(d, dtype=torch.float64) * 0.1)
        dxs.append(dx.numpy())
        xn = x + dx
        s_t = score_fn(x, t)
        m_t = x + t ** 2 * s_t
        mean, sc = bf.update_time_step(x[None], t, tn, s_t[None])
        g[f"t4_{i}_mean"], g[f"t4_{i}_newscore"] = mean.numpy()[0], sc.numpy()[0]
        m_n = xn + tn ** 2 * score_fn(xn, tn)
        # the reference test feeds the *old-sigma* mean at x as `denoiser_mean_at_x` (:362)
        bf.update_space_step(m_t[None], m_n[None], tn, x[None], xn[None])
        for n, m in zip("cov inv_cov hes inv_hes".split(), dense4(bf)):
            g[f"t4_{i}_{n}"] = m
        x = xn
    g["t4_dxs"] = np.stack(dxs)
    np.savez_compressed(os.path.join(OUT, "unit_tests.npz"), **g)
    print("unit_tests.npz:", len(g), "arrays")


# ------------------------------------------------------------------------------------------
def make_inputs(seed, shape):
    rs = np.random.RandomState(seed)
    return rs.standard_normal(shape)


def golden_dct_small(ob):
    """CovarianceHessianBFGSDCT on a 3x16x16 problem (synthetic dct_variance in a temp data_dir):
    time -> space -> space -> time -> Sigma.v, all outputs stored in full."""
    g = {}
    C, H = 3, 16
    rs = np.random.RandomState(7)
    var = np.exp(rs.uniform(np.log(1e-3), np.log(50.0), size=(C, H, H))).astype(np.float32)
    tmp = tempfile.mkdtemp()
    torch.save(torch.from_numpy(var), os.path.join(tmp, "dct_variance.pt"))
    s0 = 80.0
    bf = ob.CovarianceHessianBFGSDCT(tmp, s0 ** 2, C * H * H, use_precalculated_info=True)
    g["var"] = var
    sig = [80.0, 30.0, 9.0, 9.0, 4.0, 4.0, 1.5]
    g["sigmas"] = np.array(sig)
    gain = lambda s: var.astype(np.float64) / (var + s ** 2)
    den = lambda x, s: sfft.idctn(gain(s) * sfft.dctn(x, type=2, norm="ortho", axes=(-2, -1)), type=2, norm="ortho", axes=(-2, -1))
    x = make_inputs(11, (1, C, H, H)) * s0
    xs = [x]
    step = 0
    for i in range(1, len(sig)):
        s_prev, s = sig[i - 1], sig[i]
        x_prev = xs[-1]
        mu_prev = den(x_prev, s_prev) + 0.05 * np.tanh(x_prev)          # mildly non-Gaussian denoiser
        x_new = x_prev + (s - s_prev) * (x_prev - mu_prev) / s_prev + 0.01 * make_inputs(100 + i, x.shape)
        if s != s_prev:
            score = (mu_prev - x_prev) / s_prev ** 2
            mean, sc = bf.update_time_step(torch.from_numpy(x_prev), s_prev, s, torch.from_numpy(score))
            g[f"s{i}_mean"], g[f"s{i}_score"] = mean.numpy(), sc.numpy()
            mu_hat = mean.numpy()
        else:
            mu_hat = den(x_prev, s) + 0.05 * np.tanh(x_prev)
        mu_new = den(x_new, s) + 0.05 * np.tanh(x_new)
        if 1.0 < s < 10.0:
            bf.update_space_step(torch.from_numpy(mu_hat), torch.from_numpy(mu_new), s, torch.from_numpy(x_prev), torch.from_numpy(x_new))
        w = make_inputs(200 + i, x.shape)
        g[f"s{i}_sigv"] = bf.denoiser_cov_vector_dot(torch.from_numpy(w)).numpy()
        g[f"s{i}_k"] = np.array(bf.vectors_denoiser_cov_u.shape[-1])
        xs.append(x_new)
        step += 1
    for n, m in zip("cov inv_cov hes inv_hes".split(), dense4(bf)):
        g["final_" + n] = m[::7, ::5].copy()
    np.savez_compressed(os.path.join(OUT, "dct_small.npz"), **g)
    print("dct_small.npz:", len(g), "arrays")


def synth_image(var, seed):
    """idct2(randn * sqrt(dct_variance)).clamp(-1,1) -- the synthetic clean image of SURVEY.md 8(d) C2."""
    z = make_inputs(seed, var.shape) * np.sqrt(var.astype(np.float64))
    return np.clip(sfft.idctn(z, type=2, norm="ortho", axes=(-2, -1)), -1, 1)[None]


def golden_full(ob, cm, ms):
    """Full-size (3x256x256, real dct_variance.pt) covariance model + the four customcuda mat solvers."""
    g = {}
    data_dir = os.path.join(REF, "data", "imagenet")
    var = torch.load(os.path.join(data_dir, "dct_variance.pt")).numpy()
    s0 = 80.2878
    d = var.size
    shape = (1,) + var.shape
    bf = ob.CovarianceHessianBFGSDCT(data_dir, s0 ** 2, d, use_precalculated_info=True)
    den = lambda x, s: sfft.idctn(var.astype(np.float64) / (var + s ** 2) * sfft.dctn(x, type=2, norm="ortho", axes=(-2, -1)),
                                  type=2, norm="ortho", axes=(-2, -1)) + 0.02 * np.tanh(x)
    w = make_inputs(1, shape)
    g["sigv_k0"] = sub(bf.denoiser_cov_vector_dot(torch.from_numpy(w)).numpy())
    x_clean = synth_image(var, 5)
    x0 = x_clean + s0 * make_inputs(2, shape)
    sig = [s0, 8.463, 8.463, 5.1405, 5.1405]
    xs = [x0]
    for i in range(1, len(sig)):
        s_prev, s = sig[i - 1], sig[i]
        x_prev = xs[-1]
        mu_prev = den(x_prev, s_prev)
        x_new = x_prev + (s - s_prev) * (x_prev - mu_prev) / s_prev + 0.02 * s * make_inputs(300 + i, shape)
        if s != s_prev:
            score = (mu_prev - x_prev) / s_prev ** 2
            mean, sc = bf.update_time_step(torch.from_numpy(x_prev), s_prev, s, torch.from_numpy(score))
            g[f"s{i}_mean"], g[f"s{i}_score"] = sub(mean.numpy()), sub(sc.numpy())
            mu_hat = mean.numpy()
        else:
            mu_hat = den(x_prev, s)
        bf.update_space_step(torch.from_numpy(mu_hat), torch.from_numpy(den(x_new, s)), s, torch.from_numpy(x_prev), torch.from_numpy(x_new))
        g[f"s{i}_sigv"] = sub(bf.denoiser_cov_vector_dot(torch.from_numpy(w)).numpy())
        xs.append(x_new)
    g["k_final"] = np.array(bf.vectors_denoiser_cov_u.shape[-1])
    g["sigmas"] = np.array(sig)

    # ---- operators + mat solvers at sigma_t = 5.1405 on the k=4 state
    sigma_t = sig[-1]
    x_t = xs[-1]
    mu = den(x_t, sigma_t)
    g["mu"] = sub(mu)
    noise = make_inputs(9, shape)

    def run(name, op, y):
        mat = cm.choose_solver(name, op, torch.from_numpy(y), torch.from_numpy(mu), covariance_model=bf,
                               method="customcuda", max_rtol=1.0, sigma_t=sigma_t)
        return mat.numpy()

    import conditioning_utils.cg as rcg
    niters = {}
    orig_cg = rcg.cg

    def spy(*a, **k):
        x, info = orig_cg(*a, **k)
        niters["last"] = info["niter"]
        return x, info
    rcg.cg = spy

    for name, kw in [("gaussian_blur", dict(in_shape=shape, kernel_size=61, intensity=3.0, sigma_s=0.1, device="cpu")),
                     ("motion_blur", dict(in_shape=shape, kernel_size=61, intensity=0.5, sigma_s=0.1, device="cpu"))]:
        op = ms.get_operator(name, **kw)
        y = op.forward(torch.from_numpy(x_clean), noiseless=True).numpy() + 0.1 * noise
        op.forward(torch.from_numpy(x_clean), noiseless=True)       # sets pre_calculated
        g[f"{name}_y"] = sub(y)
        g[f"{name}_Ax"] = sub(op.forward(torch.from_numpy(x_clean), noiseless=True).numpy())
        g[f"{name}_ATy"] = sub(op.transpose(torch.from_numpy(y)).numpy())
        FB = op.pre_calculated[0].numpy()
        g[f"{name}_FB_sub"] = FB.reshape(-1)[::SUB].copy()
        g[f"{name}_mat"] = sub(run(name, op, y))
        g[f"{name}_niter"] = np.array(niters["last"])

    # inpainting: random and box masks from the numpy global RNG, exactly as the reference draws them
    for mtype, opt in [("random", dict(mask_type="random", mask_prob_range=(0.6, 0.8), image_size=256)),
                       ("box", dict(mask_type="box", mask_len_range=(64, 156), image_size=256))]:
        np.random.seed(123)
        torch.manual_seed(0)
        op = ms.get_operator("inpainting", device="cpu", sigma_s=0.1, mask_opt=opt)
        mask = op.mask.numpy()
        g[f"inpaint_{mtype}_mask_bits"] = np.packbits(mask[0, 0].astype(np.uint8))
        y = (x_clean + 0.1 * noise) * mask
        g[f"inpaint_{mtype}_mat"] = sub(run("inpainting", op, y))
        g[f"inpaint_{mtype}_niter"] = np.array(niters["last"])

    # super-resolution x4: y from the FFT operator (the solver's A), kernels[0,2]
    op = ms.get_operator("super_resolution", in_shape=shape, scale_factor=4, sigma_s=0.1, device="cpu")
    k = op.get_kernel()
    import measurement_utils.utils_sisr as sisr
    FB, FBC, F2B, _ = sisr.pre_calculate(torch.zeros(1, 3, 64, 64), k, 4)
    y = sisr.downsample(torch.fft.ifft2(FB * torch.fft.fft2(torch.from_numpy(x_clean))).real, 4).numpy()
    y = y + 0.1 * make_inputs(10, y.shape)
    op.pre_calculated = sisr.pre_calculate(torch.from_numpy(y), k, 4)
    g["sr_kernel"] = k.numpy()[0, 0]
    g["sr_y"] = y
    g["sr_mat"] = sub(run("super_resolution", op, y))
    g["sr_niter"] = np.array(niters["last"])
    rcg.cg = orig_cg
    np.savez_compressed(os.path.join(OUT, "full_size.npz"), **g)
    print("full_size.npz:", len(g), "arrays", {k: int(v) for k, v in g.items() if k.endswith("niter")})


def golden_mechanism(ob, cm, ms, max_rtol=1.0, t_steps=None, fname="mechanism.npz"):
    """BFGSOnlineUpdate (the Free Hunch mechanism) end to end on a short Heun grid with the analytic
    DCT-Gaussian denoiser (differentiable torch version, float64), gaussian_blur.
    max_rtol=1 is the reference default -- and chaotic: CG stops on `|r| <= ~1*|b|` after ~100 iterations with
    x0=b, so 1e-14 input perturbations move the outputs by tens of percent (see tests/test_oracle_golden.py).
    max_rtol=1e-6 converges every solve and pins the orchestration tightly."""
    g = {}
    data_dir = os.path.join(REF, "data", "imagenet")
    var = torch.load(os.path.join(data_dir, "dct_variance.pt")).to(torch.float64)
    shape = (1, 3, 256, 256)
    n = 256
    kk = np.arange(n)[:, None]; nn = np.arange(n)[None, :]
    Cm = np.sqrt(2.0 / n) * np.cos(np.pi * (2 * nn + 1) * kk / (2 * n)); Cm[0] /= np.sqrt(2.0)
    Cm = torch.from_numpy(Cm)

    def model(x, sigma):
        gain = var / (var + sigma ** 2)
        xh = Cm @ x.to(torch.float64) @ Cm.T
        return Cm.T @ (gain * xh) @ Cm, None

    x_clean = synth_image(var.numpy(), 5)
    op = ms.get_operator("gaussian_blur", in_shape=shape, kernel_size=61, intensity=3.0, sigma_s=0.1, device="cpu")
    y = op.forward(torch.from_numpy(x_clean), noiseless=True) + 0.1 * torch.from_numpy(make_inputs(9, shape))
    if t_steps is None:
        t_steps = np.array([80.2878, 20.9185, 8.463, 5.1405, 2.2699, 0.8932, 0.2053, 0.0335, 0.0])
    os.chdir(REF)
    mech = cm.BFGSOnlineUpdate(1.0, op, False, init_denoiser_variance=1, init_noise_variance=torch.tensor(t_steps[0] ** 2, dtype=torch.float64),
                               data_dim=3 * 256 * 256, max_vector_count=100000, data_dir=data_dir, image_base_covariance="dct_diagonal",
                               pca_component_count=0, denoiser_mean_error_threshold=0.2, use_analytical_score_time_update=True,
                               project_to_diagonal=False, space_step_update_threshold=10.0, space_step_update_lower_threshold=1.0,
                               max_rtol=max_rtol, do_space_updates=True, use_analytic_var_at_end=False, solver_type="customcuda",
                               use_rtol_func=True, diffpir_lambda=0)
    import conditioning_utils.cg as rcg
    niters = []
    orig_cg = rcg.cg

    def spy(*a, **k):
        x, info = orig_cg(*a, **k)
        niters.append(info["niter"])
        return x, info
    rcg.cg = spy
    x_next = torch.from_numpy(make_inputs(2, shape)) * t_steps[0]
    call = 0
    for i, (tc, tn) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_hat = x_next.detach().clone()
        with torch.enable_grad():
            den = mech(x_hat.requires_grad_(), model, y, torch.tensor(tc, dtype=torch.float64)).detach()
        g[f"call{call}_out"] = sub(den.numpy()); call += 1
        d_cur = (x_hat.detach() - den) / tc
        h = tn - tc
        x_prime = x_hat.detach() + h * d_cur
        if i == len(t_steps) - 2:
            x_next = x_hat.detach() + h * d_cur
        else:
            with torch.enable_grad():
                den2 = mech(x_prime.clone().requires_grad_(), model, y, torch.tensor(tn, dtype=torch.float64)).detach()
            g[f"call{call}_out"] = sub(den2.numpy()); call += 1
            d_prime = x_prime / tn - den2 / tn
            x_next = x_hat.detach() + h * (0.5 * d_cur + 0.5 * d_prime)
    rcg.cg = orig_cg
    g["t_steps"] = t_steps
    g["FB"] = op.pre_calculated[0].numpy()[0, 0]       # the reference's complex64 OTF (torch float32 FFT)
    g["niters"] = np.array(niters)
    g["final"] = sub(x_next.numpy())
    g["final_norm"] = np.array(np.linalg.norm(x_next.numpy()))
    g["k_final"] = np.array(mech.covariance_model.vectors_denoiser_cov_u.shape[-1])
    g["max_rtol"] = np.array(max_rtol)
    np.savez_compressed(os.path.join(OUT, fname), **g)
    print(fname, ": calls", call, "niters", niters, "k", int(g["k_final"]))


def golden_config(ob, cm, ms, operator, num_steps, solver, max_rtol, fname, seed=5):
    """A FULL BASELINE configuration through the real reference: BFGSOnlineUpdate inside the EDM Heun / Euler loop of
    generate_conditional.py:131-167 on the complete `num_steps` sigma grid (30-step Heun = 59 mechanism calls, 100-step
    Euler = 100), analytic DCT-Gaussian denoiser (float64 dense DCT), for gaussian_blur / motion_blur / inpainting
    (random mask, p in [0.6, 0.8]) / super_resolution x4.  Stores the per-call outputs and the final sample (subsampled),
    the CG iteration count of every call, and whatever the test needs to rebuild the same inputs (OTF, mask bits, y)."""
    sys.path.insert(0, os.path.join(os.path.dirname(OUT), ".."))
    from oracle import fh_oracle as fo
    g = {}
    data_dir = os.path.join(REF, "data", "imagenet")
    var = torch.load(os.path.join(data_dir, "dct_variance.pt")).to(torch.float64)
    shape = (1, 3, 256, 256)
    n = 256
    kk = np.arange(n)[:, None]; nn = np.arange(n)[None, :]
    Cm = np.sqrt(2.0 / n) * np.cos(np.pi * (2 * nn + 1) * kk / (2 * n)); Cm[0] /= np.sqrt(2.0)
    Cm = torch.from_numpy(Cm)

    def model(x, sigma):
        gain = var / (var + sigma ** 2)
        xh = Cm @ x.to(torch.float64) @ Cm.T
        return Cm.T @ (gain * xh) @ Cm, None

    x_clean = synth_image(var.numpy(), seed)
    os.chdir(REF)
    if operator == "inpainting":
        np.random.seed(seed)
        torch.manual_seed(seed)
        opt = dict(mask_type="random", mask_prob_range=(0.6, 0.8), image_size=256)
        op = ms.get_operator("inpainting", device="cpu", sigma_s=0.1, mask_opt=opt)      # draws the mask (numpy global RNG)
        mask = op.mask.to(torch.float64)
        g["mask_bits"] = np.packbits(mask.numpy()[0, 0].astype(np.uint8))
        y = (torch.from_numpy(x_clean) + 0.1 * torch.from_numpy(make_inputs(seed + 4, shape))) * mask
    elif operator == "super_resolution":
        op = ms.get_operator("super_resolution", in_shape=shape, scale_factor=4, sigma_s=0.1, device="cpu")
        y = op.forward(torch.from_numpy(x_clean).float(), noiseless=True).to(torch.float64)
        y = y + 0.1 * torch.from_numpy(make_inputs(seed + 4, tuple(y.shape)))
        g["y"] = y.numpy().astype(np.float32)
        y = torch.from_numpy(g["y"])
    else:
        kw = dict(intensity=3.0) if operator == "gaussian_blur" else dict(intensity=0.5)
        op = ms.get_operator(operator, in_shape=shape, kernel_size=61, sigma_s=0.1, device="cpu", **kw)
        y = op.forward(torch.from_numpy(x_clean), noiseless=True) + 0.1 * torch.from_numpy(make_inputs(seed + 4, shape))
        if operator == "motion_blur":
            g["FB"] = op.pre_calculated[0].numpy()[0, 0]
    t_steps = fo.edm_sigma_steps(num_steps)
    mech = cm.BFGSOnlineUpdate(1.0, op, False, init_denoiser_variance=1, init_noise_variance=torch.tensor(t_steps[0] ** 2, dtype=torch.float64),
                               data_dim=3 * 256 * 256, max_vector_count=100000, data_dir=data_dir, image_base_covariance="dct_diagonal",
                               pca_component_count=0, denoiser_mean_error_threshold=0.2, use_analytical_score_time_update=True,
                               project_to_diagonal=False, space_step_update_threshold=10.0, space_step_update_lower_threshold=1.0,
                               max_rtol=max_rtol, do_space_updates=True, use_analytic_var_at_end=False, solver_type="customcuda",
                               use_rtol_func=True, diffpir_lambda=0)
    import conditioning_utils.cg as rcg
    import time
    niters = []
    orig_cg = rcg.cg

    def spy(*a, **k):
        x, info = orig_cg(*a, **k)
        niters.append(info["niter"])
        return x, info
    rcg.cg = spy
    cm.cg = spy if hasattr(cm, "cg") else None
    t0 = time.time()
    x_next = torch.from_numpy(make_inputs(seed - 3, shape)) * t_steps[0]
    call = 0
    N = len(t_steps) - 1
    for i, (tc, tn) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_hat = x_next.detach().clone()
        with torch.enable_grad():
            den = mech(x_hat.requires_grad_(), model, y, torch.tensor(tc, dtype=torch.float64)).detach()
        g[f"call{call}_out"] = sub(den.numpy()).astype(np.float32) if call % 4 else sub(den.numpy()); call += 1
        d_cur = (x_hat.detach() - den) / tc
        h = tn - tc
        x_prime = x_hat.detach() + h * d_cur
        if solver == "euler" or i == N - 1:
            x_next = x_prime
        else:
            with torch.enable_grad():
                den2 = mech(x_prime.clone().requires_grad_(), model, y, torch.tensor(tn, dtype=torch.float64)).detach()
            g[f"call{call}_out"] = sub(den2.numpy()); call += 1
            d_prime = x_prime / tn - den2 / tn
            x_next = x_hat.detach() + h * (0.5 * d_cur + 0.5 * d_prime)
        print(f"  [{fname}] step {i + 1}/{N} sigma {tc:.4f} calls {call} cg {sum(niters)} t {time.time() - t0:.0f}s", flush=True)
    rcg.cg = orig_cg
    g["t_steps"] = np.asarray(t_steps)
    g["niters"] = np.array(niters)
    g["final"] = sub(x_next.numpy())
    g["final_norm"] = np.array(np.linalg.norm(x_next.numpy()))
    g["k_final"] = np.array(mech.covariance_model.vectors_denoiser_cov_u.shape[-1])
    g["max_rtol"] = np.array(max_rtol)
    g["seed"] = np.array(seed)
    np.savez_compressed(os.path.join(OUT, fname), **g)
    print(fname, ": calls", call, "cg total", sum(niters), "k", int(g["k_final"]), f"{time.time() - t0:.0f}s")


CONFIGS = {   # name -> (operator, num_steps, solver, max_rtol)
    "c1_gauss_heun30": ("gaussian_blur", 30, "heun", 1.0),
    "c2_inpaint_heun30": ("inpainting", 30, "heun", 1.0),
    "c3_sr_heun30": ("super_resolution", 30, "heun", 1.0),
    "c4_motion_euler100": ("motion_blur", 100, "euler", 1.0),
    "c1_gauss_heun30_tight": ("gaussian_blur", 30, "heun", 1e-6),
    "c2_inpaint_heun30_tight": ("inpainting", 30, "heun", 1e-6),
    "c3_sr_heun30_tight": ("super_resolution", 30, "heun", 1e-6),
    "c4_motion_euler100_tight": ("motion_blur", 100, "euler", 1e-6),
}


def golden_misc(cm):
    """rtol schedule, sigma grid (round_sigma of the iDDPM-linear wrapper), toy Gaussian mixture."""
    g = {}
    from training.openai_preconditioning import iDDPMLinearPrecond
    net = iDDPMLinearPrecond(model=None, img_resolution=256, img_channels=3)
    g["u"] = net.u.numpy()
    for N in (30, 100):
        smin, smax = max(0.002, net.sigma_min), min(80.0, net.sigma_max)
        idx = torch.arange(N, dtype=torch.float64)
        steps = (smax ** (1 / 7) + idx / (N - 1) * (smin ** (1 / 7) - smax ** (1 / 7))) ** 7   # generate_conditional.py:200
        g[f"grid{N}"] = net.round_sigma(steps).numpy()
    sig = np.concatenate([g["grid30"], [100.0, 0.05, 0.1, 80.0]])
    g["rtol_sigmas"] = sig
    g["rtol_1"] = np.array([cm.rtol_func(float(s), 1.0) for s in sig])
    g["rtol_01"] = np.array([cm.rtol_func(float(s), 0.1) for s in sig])

    # toy Gaussian mixture (config 1), 64 samples instead of 10000 to keep the fixture small
    sys.path.insert(0, os.path.join(REF, "notebooks"))
    import diffusion_gaussian_mixture_utils as gm
    weights = torch.tensor([0.3, 0.7])
    means = torch.tensor([[-1.0, -1.0], [1.0, 1.0]])
    covs = torch.tensor([[[1.0, 0.9], [0.9, 1.0]], [[1.0, -0.6], [-0.6, 1.0]]])
    cov_y = torch.tensor([[0.5, 0.0], [0.0, 0.5]])
    y = torch.tensor([-0.7, 0.4])
    torch.set_default_dtype(torch.float64)
    weights, means, covs, cov_y, y = [t.to(torch.float64) for t in (weights, means, covs, cov_y, y)]
    smin, smax = 0.001, 15.0
    schedule = lambda t: smin + (smax - smin) * t
    dschedule = lambda t: (smax - smin)
    score_fn = lambda x, t: gm.gaussian_mixture_log_gradient(x, weights, means, covs, schedule(t))
    torch.manual_seed(42)
    prior = torch.randn(64, 2) * 10.0
    prior[0, :] = torch.tensor([4.0, -4.0])
    init_cov = gm.gaussian_mixture_covariance(weights, means, covs)
    out = gm.sample_ode_with_second_order_bfgs_updates(prior, schedule, dschedule, score_fn, 100, init_cov, y, "cpu",
                                                       weights, means, covs, cov_y, rho=2.0, get_sample_paths=True)
    x, xs, ts, dcovs = out[0], out[1], out[2], out[3]
    g["gmm_prior"] = prior.numpy()
    g["gmm_init_cov"] = init_cov.numpy()
    g["gmm_final"] = x.detach().numpy()
    g["gmm_cov_50"] = dcovs[50].numpy()
    g["gmm_cov_last"] = dcovs[-1].numpy()
    # exact denoiser covariance at the final samples, sigma at the last evaluated time
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **g)
    print("misc.npz:", len(g), "arrays")


def export_assets():
    """Data assets the tests/bench need on the GPU box (the reference tree does not travel):
    the per-DCT-coefficient variance and the three PSFs, repacked into one npz."""
    var = torch.load(os.path.join(REF, "data/imagenet/dct_variance.pt")).numpy()
    gk = np.load(os.path.join(REF, "measurement_utils/kernels/gaussian_ks61_std3.0.npy"))
    mk = np.load(os.path.join(REF, "measurement_utils/kernels/motion_ks61_std0.5.npy"))
    bk = scipy.io.loadmat(os.path.join(REF, "measurement_utils/kernels/kernels_bicubicx234.mat"))["kernels"]
    np.savez_compressed(os.path.join(OUT, "assets.npz"), dct_variance=var, gaussian_ks61_std3=gk,
                        motion_ks61_std05=mk, bicubic_x2=bk[0, 0], bicubic_x3=bk[0, 1], bicubic_x4=bk[0, 2])
    print("assets.npz written")


if __name__ == "__main__":
    ob, cm, rcg, ms, sisr = import_reference()
    which = sys.argv[1:] or ["unit", "small", "misc", "assets", "full", "mech", "mech_tight"]
    torch.set_num_threads(int(os.environ.get("FH_GOLDEN_THREADS", os.cpu_count())))
    if "assets" in which: export_assets()
    if "unit" in which: golden_unit(ob)
    if "small" in which: golden_dct_small(ob)
    if "misc" in which: golden_misc(cm)
    if "full" in which: golden_full(ob, cm, ms)
    if "mech" in which: golden_mechanism(ob, cm, ms)
    for name, (opn, ns, solver, mr) in CONFIGS.items():
        if name in which:
            golden_config(ob, cm, ms, opn, ns, solver, mr, f"cfg_{name}.npz")
    if "mech_tight" in which:
        golden_mechanism(ob, cm, ms, max_rtol=1e-6, t_steps=np.array([80.2878, 8.463, 5.1405, 2.2699, 0.2053, 0.0]),
                         fname="mechanism_tight.npz")
