"""Pins oracle/fh_oracle.py against outputs of the REAL reference (tests/golden/*.npz, produced by
tests/golden/make_golden.py in the build container).  CPU only."""
import os
import sys

import numpy as np
import pytest
import scipy.fft as sfft

from conftest import load_golden, rel_l2

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fh_oracle as fo  # noqa: E402

SUB = 97
NAMES = "cov inv_cov hes inv_hes".split()


def sub(a):
    return np.asarray(a, np.float64).reshape(-1)[::SUB]


def make_inputs(seed, shape):
    return np.random.RandomState(seed).standard_normal(shape)


def dense4(c):
    return [m.real for m in c.get_dense_matrices()]


def check4(c, g, prefix, tol=1e-9):
    for n, m in zip(NAMES, dense4(c)):
        ref = g[prefix + n]
        # inverse Hessians carry 1/(tiny) entries; compare relative to the matrix norm
        assert rel_l2(m, ref) < tol, (prefix + n, rel_l2(m, ref))


# ------------------------------------------------------------------ reference unit tests (F3-adapted)
def test_unit_time_update():
    g = load_golden("unit_tests.npz")
    x, score = g["t1_x"], g["t1_score"]
    c = fo.CovarianceOracle(1, 20.0 ** 2, 5)
    mean, sc = c.update_time_step(x[None], 20.0, 18.0, score[None])
    assert rel_l2(mean[0], g["t1_mean"]) < 1e-12
    assert rel_l2(sc[0], g["t1_newscore"]) < 1e-12
    check4(c, g, "t1_")
    # low-rank class == dense form (what the reference test prints)
    eye = np.eye(5)
    hes = (eye / 400 - eye) / 400
    out = fo.dense_update_covariance(x, eye, eye.copy(), hes, np.linalg.inv(hes), score, 20.0, 18.0)
    for n, m in zip("cov inv_cov hes inv_hes newscore mean".split(), out):
        assert rel_l2(m, g["t1_dense_" + n]) < 1e-10, n
    for m, n in zip(dense4(c), NAMES):
        assert rel_l2(m, g["t1_dense_" + n]) < 1e-8, n


@pytest.mark.parametrize("pairs", [1, 2, 4, 8])
def test_unit_time_update_with_u_v(pairs):
    g = load_golden("unit_tests.npz")
    p = f"t2_{pairs}_"
    x = g[p + "x"]                                    # float32, as in the reference test
    score = -x / np.float32(80.0 ** 2)                # float32 division, like torch
    c = fo.CovarianceOracle(1, 80.0 ** 2, 15)
    c.cov_u = g[p + "U"].astype(np.complex128)
    c.cov_v = g[p + "V"].astype(np.complex128)
    c.set_others_corresponding_to_current_denoiser_cov(80.0)
    check4(c, g, p + "pre_", tol=1e-7)
    mean, sc = c.update_time_step(x[None], 80.0, 79.0, score[None])
    assert rel_l2(mean[0], g[p + "mean"]) < 1e-7
    assert rel_l2(sc[0], g[p + "newscore"]) < 1e-7
    check4(c, g, p, tol=1e-6)


def _score_fn(x, t):
    return 0.7 * (-x / t ** 2) + 0.3 * (-0.5 * (x - 1.0) / t ** 2)


def test_unit_bfgs_update():
    g = load_golden("unit_tests.npz")
    d, t = 15, 50.0
    x = g["t3_x0"].astype(np.float32)
    c = fo.CovarianceOracle(1, t ** 2, d)
    den = np.eye(d)[None]
    inv = np.eye(d)[None]
    for step in range(10):
        dx = g["t3_dxs"][step].astype(np.float32)
        xn = x + dx
        m0 = (x + np.float32(t ** 2) * _score_fn(x, np.float32(t))).astype(np.float32)
        m1 = (xn + np.float32(t ** 2) * _score_fn(xn, np.float32(t))).astype(np.float32)
        c.update_space_step(m0[None], m1[None], t, x[None], xn[None])
        # float32 inputs reach the reference class as float32 -> complex128; inputs recomputed here in
        # float32 may differ from torch's in the last ulp, hence the looser bound
        check4(c, g, f"t3_{step}_", tol=2e-4)
        den, inv, _, _ = fo.dense_update_bfgs(den, inv, m0[None].astype(np.float64), m1[None].astype(np.float64), t,
                                              dx[None].astype(np.float64))
        assert rel_l2(den[0], dense4(c)[0]) < 1e-6
        assert rel_l2(inv[0], dense4(c)[1]) < 1e-5
        x = xn


def test_unit_time_and_space():
    g = load_golden("unit_tests.npz")
    d = 5
    ts = g["t4_ts"]
    x = g["t4_x0"].astype(np.float64)
    c = fo.CovarianceOracle(1, ts[0] ** 2, d)
    for i in range(len(ts) - 1):
        t, tn = ts[i], ts[i + 1]
        dx = g["t4_dxs"][i]
        xn = x + dx
        s_t = _score_fn(x, t)
        m_t = x + t ** 2 * s_t
        mean, sc = c.update_time_step(x[None], t, tn, s_t[None])
        assert rel_l2(mean[0], g[f"t4_{i}_mean"]) < 1e-6, i
        assert rel_l2(sc[0], g[f"t4_{i}_newscore"]) < 1e-6, i
        m_n = xn + tn ** 2 * _score_fn(xn, tn)
        c.update_space_step(m_t[None], m_n[None], tn, x[None], xn[None])
        for n, m in zip(NAMES, dense4(c)):
            # the Hessian families pass through 1/(~0) diagonals; the reference itself only holds ~1e-4 there
            tol = 1e-6 if n in ("cov", "inv_cov") else 1e-3
            assert rel_l2(m, g[f"t4_{i}_{n}"]) < tol, (i, n, rel_l2(m, g[f"t4_{i}_{n}"]))
        x = xn


# ------------------------------------------------------------------ DCT class, small problem
def test_dct_small_sequence():
    g = load_golden("dct_small.npz")
    var = g["var"]
    C, H = 3, 16
    sig = list(g["sigmas"])
    c = fo.CovarianceOracleDCT(var, sig[0] ** 2, C * H * H)
    gain = lambda s: var.astype(np.float64) / (var + s ** 2)
    den = lambda x, s: fo.idct2(gain(s) * fo.dct2(x))
    x = make_inputs(11, (1, C, H, H)) * sig[0]
    xs = [x]
    for i in range(1, len(sig)):
        s_prev, s = sig[i - 1], sig[i]
        x_prev = xs[-1]
        mu_prev = den(x_prev, s_prev) + 0.05 * np.tanh(x_prev)
        x_new = x_prev + (s - s_prev) * (x_prev - mu_prev) / s_prev + 0.01 * make_inputs(100 + i, x.shape)
        if s != s_prev:
            score = (mu_prev - x_prev) / s_prev ** 2
            mean, sc = c.update_time_step(x_prev, s_prev, s, score)
            assert rel_l2(mean, g[f"s{i}_mean"]) < 1e-8, i
            assert rel_l2(sc, g[f"s{i}_score"]) < 1e-8, i
            mu_hat = mean
        else:
            mu_hat = den(x_prev, s) + 0.05 * np.tanh(x_prev)
        mu_new = den(x_new, s) + 0.05 * np.tanh(x_new)
        if 1.0 < s < 10.0:
            c.update_space_step(mu_hat, mu_new, s, x_prev, x_new)
        w = make_inputs(200 + i, x.shape)
        assert rel_l2(c.denoiser_cov_vector_dot(w), g[f"s{i}_sigv"]) < 1e-8, i
        assert c.cov_u.shape[-1] == int(g[f"s{i}_k"])
        xs.append(x_new)
    for n, m in zip(NAMES, dense4(c)):
        tol = 1e-7 if n in ("cov", "inv_cov") else 1e-4
        assert rel_l2(m[::7, ::5], g["final_" + n]) < tol, n


# ------------------------------------------------------------------ misc: rtol, sigma grid, toy mixture
def test_rtol_and_sigma_grid():
    g = load_golden("misc.npz")
    u = fo.iddpm_linear_u()
    # float32 cumprod near alpha_bar ~ 1 amplifies evaluation-order differences between numpy and torch:
    # the numpy restatement agrees to 1e-4; the rounding test below therefore uses the reference's own u.
    assert np.allclose(u, g["u"], rtol=1e-4, atol=0)
    # the rounded grid must land on the same u entries as the reference's
    for N in (30, 100):
        grid = fo.edm_sigma_steps(N, u=g["u"])[:-1]
        assert np.array_equal(grid, g[f"grid{N}"].astype(np.float64)), N
    own = fo.edm_sigma_steps(30)
    assert np.allclose(own[:-1], g["grid30"], rtol=1e-4)
    for s, r1, r01 in zip(g["rtol_sigmas"], g["rtol_1"], g["rtol_01"]):
        assert fo.rtol_func(float(s), 1.0) == pytest.approx(r1, rel=1e-12)
        assert fo.rtol_func(float(s), 0.1) == pytest.approx(r01, rel=1e-12)


def test_gmm_toy_config1():
    g = load_golden("misc.npz")
    # the reference run built these as float32 tensors and then widened them (make_golden.py)
    f32 = lambda a: np.array(a, np.float32).astype(np.float64)
    weights = f32([0.3, 0.7])
    means = f32([[-1.0, -1.0], [1.0, 1.0]])
    covs = f32([[[1.0, 0.9], [0.9, 1.0]], [[1.0, -0.6], [-0.6, 1.0]]])
    cov_y = 0.5 * np.eye(2)
    y = f32([-0.7, 0.4])
    smin, smax = 0.001, 15.0
    init_cov = fo.gmm_covariance(weights, means, covs)
    assert rel_l2(init_cov, g["gmm_init_cov"]) < 1e-12
    x, dcovs = fo.gmm_sample_bfgs(g["gmm_prior"], lambda t: smin + (smax - smin) * t, lambda t: smax - smin, 100,
                                  init_cov, y, weights, means, covs, cov_y, rho=2.0)
    # per-sample comparison: chaotic amplification over 100 BFGS steps is bounded but not tiny
    assert rel_l2(dcovs[50], g["gmm_cov_50"]) < 1e-6
    assert rel_l2(x, g["gmm_final"]) < 1e-5
    assert rel_l2(dcovs[-1], g["gmm_cov_last"]) < 1e-4


# ------------------------------------------------------------------ full size
def _full_state(assets):
    var = assets["dct_variance"]
    s0 = 80.2878
    shape = (1,) + var.shape
    c = fo.CovarianceOracleDCT(var, s0 ** 2, var.size)
    den = lambda x, s: fo.idct2(var.astype(np.float64) / (var + s ** 2) * fo.dct2(x)) + 0.02 * np.tanh(x)
    return var, s0, shape, c, den


def synth_image(var, seed):
    z = make_inputs(seed, var.shape) * np.sqrt(var.astype(np.float64))
    return np.clip(fo.idct2(z), -1, 1)[None]


@pytest.fixture(scope="module")
def full_run(assets):
    g = load_golden("full_size.npz")
    var, s0, shape, c, den = _full_state(assets)
    out = {}
    w = make_inputs(1, shape)
    out["sigv_k0"] = sub(c.denoiser_cov_vector_dot(w))
    x_clean = synth_image(var, 5)
    x0 = x_clean + s0 * make_inputs(2, shape)
    sig = list(g["sigmas"])
    xs = [x0]
    for i in range(1, len(sig)):
        s_prev, s = sig[i - 1], sig[i]
        x_prev = xs[-1]
        mu_prev = den(x_prev, s_prev)
        x_new = x_prev + (s - s_prev) * (x_prev - mu_prev) / s_prev + 0.02 * s * make_inputs(300 + i, shape)
        if s != s_prev:
            score = (mu_prev - x_prev) / s_prev ** 2
            mean, sc = c.update_time_step(x_prev, s_prev, s, score)
            out[f"s{i}_mean"], out[f"s{i}_score"] = sub(mean), sub(sc)
            mu_hat = mean
        else:
            mu_hat = den(x_prev, s)
        c.update_space_step(mu_hat, den(x_new, s), s, x_prev, x_new)
        out[f"s{i}_sigv"] = sub(c.denoiser_cov_vector_dot(w))
        xs.append(x_new)
    return dict(g=g, out=out, c=c, den=den, xs=xs, x_clean=x_clean, shape=shape, sig=sig, assets=assets)


def test_full_covariance_sequence(full_run):
    g, out = full_run["g"], full_run["out"]
    assert full_run["c"].cov_u.shape[-1] == int(g["k_final"])
    for k, v in out.items():
        assert rel_l2(v, g[k]) < 1e-7, (k, rel_l2(v, g[k]))


def _solve(full_run, op, y):
    sigma_t = full_run["sig"][-1]
    mu = full_run["den"](full_run["xs"][-1], sigma_t)
    return fo.mat_solver(op, y, mu, full_run["c"], sigma_t, max_rtol=1.0)


@pytest.mark.parametrize("name,key", [("gaussian_blur", "gaussian_ks61_std3"), ("motion_blur", "motion_ks61_std05")])
def test_full_blur_solvers(full_run, name, key):
    g, a = full_run["g"], full_run["assets"]
    cls = fo.GaussianBlurOracle if name == "gaussian_blur" else fo.MotionBlurOracle
    op = cls(a[key], 0.1)
    assert rel_l2(op.FB.reshape(-1)[::SUB], g[f"{name}_FB_sub"]) < 2e-6      # complex64 FFT rounding
    x_clean = full_run["x_clean"]
    noise = make_inputs(9, full_run["shape"])
    y = op.A(x_clean) + 0.1 * noise
    assert rel_l2(sub(op.A(x_clean)), g[f"{name}_Ax"]) < 1e-6
    assert rel_l2(sub(y), g[f"{name}_y"]) < 1e-6
    assert rel_l2(sub(op.AT(y)), g[f"{name}_ATy"]) < 1e-6
    mat, info = _solve(full_run, op, y)
    assert info["niter"] == int(g[f"{name}_niter"])
    assert rel_l2(sub(mat), g[f"{name}_mat"]) < 1e-5


@pytest.mark.parametrize("mtype", ["random", "box"])
def test_full_inpainting_solver(full_run, mtype):
    g = full_run["g"]
    np.random.seed(123)
    if mtype == "random":
        mask = fo.mask_random(256, (0.6, 0.8))
    else:
        mask = fo.mask_box(256, (64, 156))
    # bit-exact mask
    assert np.array_equal(np.packbits(mask[0, 0].astype(np.uint8)), g[f"inpaint_{mtype}_mask_bits"])
    assert np.array_equal(mask[0, 0], mask[0, 1]) and np.array_equal(mask[0, 0], mask[0, 2])
    op = fo.InpaintingOracle(mask, 0.1)
    y = (full_run["x_clean"] + 0.1 * make_inputs(9, full_run["shape"])) * mask
    mat, info = _solve(full_run, op, y)
    assert info["niter"] == int(g[f"inpaint_{mtype}_niter"])
    assert rel_l2(sub(mat), g[f"inpaint_{mtype}_mat"]) < 1e-6


def test_full_sr_solver(full_run):
    g, a = full_run["g"], full_run["assets"]
    assert np.array_equal(a["bicubic_x4"], g["sr_kernel"])
    op = fo.SuperResolutionOracle(a["bicubic_x4"], 0.1, scale_factor=4)
    mat, info = _solve(full_run, op, g["sr_y"])
    assert info["niter"] == int(g["sr_niter"])
    assert rel_l2(sub(mat), g["sr_mat"]) < 1e-5


# ------------------------------------------------------------------ mechanism end to end
def _run_mechanism(assets, g, noise_eps=0.0):
    var = assets["dct_variance"]
    shape = (1,) + var.shape
    t_steps = g["t_steps"]
    op = fo.GaussianBlurOracle(assets["gaussian_ks61_std3"], 0.1)
    # use the reference's own complex64 OTF bits (torch's float32 FFT differs from scipy's in the last ulp)
    assert rel_l2(op.FB, g["FB"][None, None]) < 2e-6
    op.FB = g["FB"][None, None]
    op.FBC = np.conj(op.FB)
    x_clean = synth_image(var, 5)
    y = op.A(x_clean) + 0.1 * make_inputs(9, shape)
    cov = fo.CovarianceOracleDCT(var, t_steps[0] ** 2, var.size)
    mech = fo.FreeHunchOracle(op, cov, max_rtol=float(g["max_rtol"]) if "max_rtol" in g.files else 1.0)
    model = fo.analytic_dct_gaussian_denoiser(var)
    outs = []

    def spy(x_t, mdl, yy, s):
        o = mech(x_t, mdl, yy, s)
        outs.append(sub(o))
        return o
    noise = make_inputs(2, shape)
    if noise_eps:
        noise = noise * (1 + noise_eps * make_inputs(77, shape))
    x = fo.conditional_sampler(spy, model, y, noise, t_steps, solver="heun")
    return mech, cov, outs, x


def test_mechanism_end_to_end_converged(assets):
    """max_rtol=1e-6: every CG solve converges, the trajectory is well defined, and the oracle must follow
    the reference call by call (time/space-update orchestration, fallback rule, Heun loop)."""
    g = load_golden("mechanism_tight.npz")
    mech, cov, outs, x = _run_mechanism(assets, g)
    ref_n = np.array(g["niters"])
    got_n = np.array([r["niter"] for r in mech.log])
    assert np.all(np.abs(got_n - ref_n) <= np.maximum(2, 0.02 * ref_n)), (got_n, ref_n)
    assert cov.cov_u.shape[-1] == int(g["k_final"])
    for i, o in enumerate(outs):
        assert rel_l2(o, g[f"call{i}_out"]) < 1e-5, (i, rel_l2(o, g[f"call{i}_out"]))
    assert rel_l2(sub(x), g["final"]) < 1e-5


def test_mechanism_default_rtol_is_chaotic(assets):
    """max_rtol=1 (reference default).  At sigma~80 the stop rule |r| <= ~1*|b| is met after ~30-120 CG
    iterations started from x0=b; the iterate at the stop is NOT a converged quantity.  This test records the
    fact the parity tolerances of the GPU tests are built on: the first call agrees with the reference to ~1e-4,
    later high-sigma calls differ at the 10-60 % level between two float64 implementations (reference vs this
    oracle, or this oracle vs itself with a 1e-14 relative input perturbation), while the iteration counts stay
    within a few percent."""
    g = load_golden("mechanism.npz")
    mech, cov, outs, x = _run_mechanism(assets, g)
    ref_n = np.array(g["niters"])
    got_n = np.array([r["niter"] for r in mech.log])
    assert np.all(np.abs(got_n - ref_n) <= np.maximum(3, 0.05 * ref_n)), (got_n, ref_n)
    assert cov.cov_u.shape[-1] == int(g["k_final"])
    assert rel_l2(outs[0], g["call0_out"]) < 1e-3
    assert rel_l2(sub(x), g["final"]) < 5e-2
