"""CPU emulation (tools/cuda_emu.h, all 128 CTAs of the cooperative grid as fibers) of the float64 resident CG kernel
fh_k_cg_res64 (free_hunch_b200/csrc/fh_cg_res64.cuh) against the float64 oracle: with float64 in and out the only
differences left are the float32 STORAGE of delta and Z, so the bar is 1e-6 instead of the 2e-5 of the float32 kernels.
Covers the three instantiations (tables in shared memory; tables through the read-only path; rows of Z beyond the resident
capacity re-read from global memory) by shrinking the capacity through FH_EMU_R64_CAP, and the Sigma.v entry point that
runs one pass of the same kernel."""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import rel_l2

pytestmark = pytest.mark.skipif(H.HAS_CUDA, reason="CPU emulation is for boxes without a GPU")

from oracle import fh_oracle as fo  # noqa: E402


@pytest.fixture(scope="module")
def E():
    return H.use_emulation()


SMEM = {"tables in shared memory": None, "tables through L1": 34064 + 4 * 6144 + 100, "rows from L2": 34064 + 2 * 6144 + 100}


def _state(E, assets, variant, monkeypatch):
    if SMEM[variant]:
        monkeypatch.setenv("FH_EMU_R64_SMEM", str(SMEM[variant]))
    rs = np.random.RandomState(5)
    var = assets["dct_variance"]
    shape = (1,) + var.shape
    s0, s1 = 80.2878, 8.463
    ctx = E.Context("cpu", max_batch=1)
    st = E.CovState(ctx, 1, E.BASIS_DCT)
    st.init(torch.tensor(var), s0 ** 2)
    ora = fo.CovarianceOracleDCT(var, s0 ** 2, var.size)
    den = fo.analytic_dct_gaussian_denoiser(var)
    x = rs.standard_normal(shape) * s0
    mu = den(x, s0)[0] + 0.02 * np.tanh(x)
    score = (mu - x) / s0 ** 2
    st.time_update(torch.tensor(x), torch.tensor(score), s0, s1)
    m_cpu, _ = ora.update_time_step(x, s0, s1, score)
    for _ in range(2):
        xn = x + (s1 - s0) * (x - mu) / s0 + 0.3 * rs.standard_normal(shape)
        mun = den(xn, s1)[0] + 0.02 * np.tanh(xn)
        st.space_update(torch.tensor(m_cpu), torch.tensor(mun), s1, torch.tensor(x), torch.tensor(xn))
        ora.update_space_step(m_cpu, mun, s1, x, xn)
        x, m_cpu = xn, mun
    assert st.rank == 4
    return ctx, st, ora, rs, shape, s1


@pytest.mark.parametrize("variant", list(SMEM))     # rank 4; the shared-memory size picks the instantiation
def test_res64_solve_and_cov_dot(E, assets, variant, monkeypatch):
    ctx, st, ora, rs, shape, s1 = _state(E, assets, variant, monkeypatch)
    cap = 32 if SMEM[variant] is None else 0
    xc = np.clip(rs.standard_normal(shape), -1, 1)
    mu0 = 0.3 * xc
    sy2 = float(np.float32(0.1) * np.float32(0.1))
    mask = (rs.uniform(size=(256, 256)) > 0.7).astype(np.float32)
    # Sigma.v, float64 in and out
    w = rs.standard_normal(shape)
    sv = st.cov_dot(torch.tensor(w))
    assert sv.dtype == torch.float64
    assert rel_l2(sv, ora.denoiser_cov_vector_dot(w)) < 1e-6
    cases = [("inpainting", E.Operator.inpaint(ctx, mask), fo.InpaintingOracle(np.broadcast_to(mask, shape), 0.1))]
    if cap == 32:
        cases.append(("gaussian", E.Operator.conv(ctx, assets["gaussian_ks61_std3"], 1),
                      fo.GaussianBlurOracle(assets["gaussian_ks61_std3"], 0.1, otf_dtype=np.complex128)))
    for name, op, oo in cases:
        y = (xc + 0.1 * rs.standard_normal(shape)) * mask if name == "inpainting" else oo.A(xc) + 0.1 * rs.standard_normal(shape)
        y = y.astype(np.float32).astype(np.float64)                      # the library takes y as float32
        mat, info = E.cg_solve(st, op, torch.tensor(y), torch.tensor(mu0), sy2, 1.0, maxiter=50, fixed_iters=3)
        assert mat.dtype == torch.float64 and ctx.cg_last_mode == E.CG_RESIDENT
        mat_cpu, io = fo.mat_solver(oo, y, mu0, ora, s1, fixed_iters=3)
        assert info["niter"][0] == 3, name
        # gaussian: the library's taps are the rank-1 factors of the float32 PSF, the oracle's OTF is its FFT (1e-7 apart)
        assert rel_l2(mat, mat_cpu) < (1e-6 if name == "inpainting" else 5e-6), (name, cap, rel_l2(mat, mat_cpu))
        assert abs(info["residual_norm"][0] - io["residual_norm"]) <= 1e-6 * io["residual_norm"]
        if name == "inpainting" and cap == 32:
            # the residual test: a tolerance the float64 oracle first meets at iteration k <= 8 (the residual of CG started
            # from x0 = b is not monotone) stops the kernel at the same k
            b = mask * y - mask * mu0
            res = [fo.mat_solver(oo, y, mu0, ora, s1, fixed_iters=k)[1]["residual_norm"] for k in range(1, 9)]
            k = 1 + int(np.argmin(res))                    # the first iteration that reaches the smallest residual of the eight
            above = min(res[:k - 1]) if k > 1 else 2.0 * res[0]
            rtol = 0.5 * (res[k - 1] + above) / np.linalg.norm(b)
            mat, info = E.cg_solve(st, op, torch.tensor(y), torch.tensor(mu0), sy2, rtol, maxiter=200)
            assert info["niter"][0] == k and info["optimal"][0], (info, k, res)
            matk, _ = fo.mat_solver(oo, y, mu0, ora, s1, fixed_iters=k)
            assert rel_l2(mat, matk) < 1e-6

@pytest.mark.parametrize("variant", ["tables in shared memory", "tables through L1"])
def test_res64_zt_mode(E, assets, variant, monkeypatch):
    """Zt mode of fh_k_cg_res64 (FH_R64_ZT=1 forces it; by default the previous solve's iteration count decides): the
    resident rows are replaced by column-IDCT(delta z_i) in the prologue, the dots are taken before the forward column
    transform and the correction is added after the inverse one -- same iterates as the plain mode and the oracle."""
    monkeypatch.setenv("FH_R64_ZT", "1")
    test_res64_solve_and_cov_dot(E, assets, variant, monkeypatch)


def test_res64_zt_mode_follows_the_previous_solve(E, assets, monkeypatch):
    ctx, st, ora, rs, shape, s1 = _state(E, assets, "tables in shared memory", monkeypatch)
    xc = np.clip(rs.standard_normal(shape), -1, 1)
    mu0 = 0.3 * xc
    sy2 = float(np.float32(0.1) * np.float32(0.1))
    mask = (rs.uniform(size=(256, 256)) > 0.7).astype(np.float32)
    op = E.Operator.inpaint(ctx, mask)
    oo = fo.InpaintingOracle(np.broadcast_to(mask, shape), 0.1)
    y = ((xc + 0.1 * rs.standard_normal(shape)) * mask).astype(np.float32).astype(np.float64)
    outs = []
    for it in (7, 4, 4):         # 7 iterations -> the next solve runs in Zt mode (>= 6); that one reports 4 -> the third is plain again
        mat, info = E.cg_solve(st, op, torch.tensor(y), torch.tensor(mu0), sy2, 1.0, maxiter=50, fixed_iters=it)
        assert info["niter"][0] == it
        outs.append(mat)
    ref, _ = fo.mat_solver(oo, y, mu0, ora, s1, fixed_iters=4)
    assert rel_l2(outs[1], ref) < 1e-6 and rel_l2(outs[2], ref) < 1e-6
    assert not torch.equal(outs[1], outs[2])                 # two different arithmetic paths (same operator to float32 storage rounding)
    assert rel_l2(outs[1], outs[2]) < 1e-6


@pytest.mark.parametrize("which", ["motion_blur", "super_resolution", "motion_blur_zt"])
def test_res64_non_separable_operators(E, assets, which, monkeypatch):
    """The gathers of the operators that do not fold into the DCT passes (fh_r64_gather_conv: the 217-tap motion PSF as
    column runs; fh_r64_sr_gather_at / _a: 25 x 25 bicubic PSF with x4 decimation), inside the resident float64 solve:
    two CG iterations at rank 4 against the float64 oracle's FFT operators."""
    if which.endswith("_zt"):
        monkeypatch.setenv("FH_R64_ZT", "1")
        which = which[:-3]
    ctx, st, ora, rs, shape, s1 = _state(E, assets, "tables in shared memory", monkeypatch)
    xc = np.clip(rs.standard_normal(shape), -1, 1)
    mu0 = 0.3 * xc
    sy2 = float(np.float32(0.1) * np.float32(0.1))
    if which == "motion_blur":
        op = E.Operator.conv(ctx, assets["motion_ks61_std05"], 1)
        oo = fo.MotionBlurOracle(assets["motion_ks61_std05"], 0.1, otf_dtype=np.complex128)
    else:
        op = E.Operator.conv(ctx, assets["bicubic_x4"], 4)
        oo = fo.SuperResolutionOracle(assets["bicubic_x4"], 0.1, 4, otf_dtype=np.complex128)
    ay = oo.A(xc)
    y = (ay + 0.1 * rs.standard_normal(ay.shape)).astype(np.float32).astype(np.float64)
    ctx.set_cg_mode(E.CG_RESIDENT)
    try:
        mat, info = E.cg_solve(st, op, torch.tensor(y), torch.tensor(mu0), sy2, 1.0, maxiter=50, fixed_iters=2)
    finally:
        ctx.set_cg_mode(E.CG_AUTO)
    assert mat.dtype == torch.float64 and ctx.cg_last_mode == E.CG_RESIDENT and info["niter"][0] == 2
    mat_cpu, io = fo.mat_solver(oo, y, mu0, ora, s1, fixed_iters=2)
    assert rel_l2(mat, mat_cpu) < 1e-6, (which, rel_l2(mat, mat_cpu))
    assert abs(info["residual_norm"][0] - io["residual_norm"]) <= 1e-6 * io["residual_norm"]


@pytest.mark.parametrize("variant", ["tables in shared memory", "rows from L2"])
def test_fused_updates_vs_oracle(E, assets, variant, monkeypatch):
    """fh_k_upd64 (fh_upd64.cuh): the time and space updates as one cooperative launch each, float64 transfer.  Mean and score
    of every time update against the oracle (update_time_step, online_update_bfgs.py:153-192) at ranks 0, 2 and 4, Sigma.v
    after every space update, and two time updates in a row at small sigma, where the reference's float32-rounded shift of
    diag(H^-1) (:172) makes the Hessian-side variance drift from the Sigma-side one (second core CTA)."""
    if SMEM[variant]:
        monkeypatch.setenv("FH_EMU_R64_SMEM", str(SMEM[variant]))
    rs = np.random.RandomState(11)
    var = assets["dct_variance"]
    shape = (1,) + var.shape
    sig = [80.2878, 8.463, 1.31, 0.052, 0.0211, 0.0093]
    ctx = E.Context("cpu", max_batch=1)
    st = E.CovState(ctx, 1, E.BASIS_DCT)
    st.init(torch.tensor(var), sig[0] ** 2)
    ora = fo.CovarianceOracleDCT(var, sig[0] ** 2, var.size)
    den0 = fo.analytic_dct_gaussian_denoiser(var)
    den = lambda x, s: den0(x, s)[0] + 0.02 * np.tanh(x)
    w = rs.standard_normal(shape)
    x = rs.standard_normal(shape) * sig[0]
    n0 = E.launch_count()
    for i in range(1, len(sig)):
        s_prev, s = sig[i - 1], sig[i]
        mu_prev = den(x, s_prev)
        if i % 2:
            mean, sc = st.time_update(torch.tensor(x), torch.tensor((mu_prev - x) / s_prev ** 2), s_prev, s)
        else:
            mean, sc = st.time_update_from_mean(torch.tensor(x), torch.tensor(mu_prev), s_prev, s)
        mo, so = ora.update_time_step(x, s_prev, s, (mu_prev - x) / s_prev ** 2)
        assert rel_l2(mean, mo) < 1e-6 and rel_l2(sc, so) < 1e-6, (i, rel_l2(mean, mo), rel_l2(sc, so))
        if i >= 4:
            continue                                     # time updates only: var2 accumulates float32-rounded shifts
        xn = x + (s - s_prev) * (x - mu_prev) / s_prev + 0.02 * s * rs.standard_normal(shape)
        mun = den(xn, s)
        st.space_update(torch.tensor(mo), torch.tensor(mun), s, torch.tensor(x), torch.tensor(xn))
        ora.update_space_step(mo, mun, s, x, xn)
        assert st.rank == 2 * i
        e = rel_l2(st.cov_dot(torch.tensor(w)), ora.denoiser_cov_vector_dot(w))
        assert e < 1e-6, (i, e)
        x = xn
    # one launch (+ the barrier counter's memset, not a launch) per update; 3 Sigma.v launches
    assert E.launch_count() - n0 == 5 + 3 + 3, E.launch_count() - n0
    assert rel_l2(st.cov_dot(torch.tensor(w)), ora.denoiser_cov_vector_dot(w)) < 1e-6
