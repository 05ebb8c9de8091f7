"""CPU tests of the host logic and of the kernels' index/algebra through the emulation build
(tests/emu/build_emu.sh compiles free_hunch_b200/csrc with tools/cuda_emu.h).  Small sizes; the parity tests
proper are the -m gpu tests in test_gpu_parity.py.  Skipped on a GPU box (there the real library is loaded)."""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import rel_l2, make_inputs

pytestmark = pytest.mark.skipif(H.HAS_CUDA, reason="emulation tests run on CPU-only boxes")

from oracle import fh_oracle as fo  # noqa: E402


@pytest.fixture(scope="module")
def E():
    from free_hunch_b200 import _lib, engine
    saved = (_lib._lib, engine._check_device)
    yield H.use_emulation()
    _lib._lib, engine._check_device = saved


def test_emulation_is_the_test_harness_choice(E):
    from free_hunch_b200 import _lib
    assert _lib.lib()._name.endswith("libfreehunch_emu.so")


@pytest.mark.parametrize("rows", [8, 4, 2])
def test_dct_matches_scipy(E, rows, monkeypatch):
    """All three strip-height instantiations of the DCT pass kernels (the library picks by batch size on a GPU)."""
    monkeypatch.setenv("FH_EMU_ROWS", str(rows))
    ctx = E.Context("cpu", max_batch=1, channels=1, height=256, width=256)
    x = torch.from_numpy(make_inputs(0, (1, 1, 256, 256))).float()
    X = ctx.dct2(x)
    assert rel_l2(X, fo.dct2(x.numpy().astype(np.float64))) < 1e-6
    if rows == 8:
        assert rel_l2(ctx.dct2(X, inverse=True), x) < 1e-6
        assert rel_l2(ctx.dct2(x, inverse=True), fo.idct2(x.numpy().astype(np.float64))) < 1e-6


def _score_fn(x, t):
    return 0.7 * (-x / t ** 2) + 0.3 * (-0.5 * (x - 1.0) / t ** 2)


def test_reference_unit_test_time_and_space_identity_basis(E):
    case_time_and_space_identity_basis("cpu")


def case_time_and_space_identity_basis(device):
    """online_update_tests.py:274-374 (F3-adapted) through the product classes, d = 5.  Shared with tests/test_gpu_unit.py."""
    from free_hunch_b200.covariance import CovarianceHessianBFGS
    g = np.load(H.ROOT + "/tests/golden/unit_tests.npz")
    ts, d = g["t4_ts"], 5
    bf = CovarianceHessianBFGS(init_denoiser_variance=1, init_noise_variance=ts[0] ** 2, data_dim=d, device=device)
    x = torch.from_numpy(g["t4_x0"].astype(np.float64))
    for i in range(8):
        t, tn = ts[i], ts[i + 1]
        xn = x + torch.from_numpy(g["t4_dxs"][i])
        s_t = _score_fn(x, t)
        m_t = x + t ** 2 * s_t
        mean, sc = bf.update_time_step(x[None], t, tn, s_t[None])
        assert mean.dtype == torch.float64 and mean.shape == (1, d)
        assert rel_l2(mean[0], g[f"t4_{i}_mean"]) < 5e-4, i
        m_n = xn + tn ** 2 * _score_fn(xn, tn)
        bf.update_space_step(m_t[None], m_n[None], tn, x[None], xn[None])
        cov, inv_cov, hes, inv_hes = bf.get_dense_matrices()
        assert rel_l2(cov, g[f"t4_{i}_cov"]) < 1e-4, i
        assert rel_l2(inv_cov, g[f"t4_{i}_inv_cov"]) < 1e-4, i
        x = xn
    assert bf.vector_count == 8
    with pytest.raises(AssertionError, match="Batch size must be 1"):
        bf.update_time_step(x, 1.0, 0.5, x)          # un-batched input: same assertion text as the reference


@pytest.mark.parametrize("pairs", [1, 4])
def test_reference_unit_test_preseeded_vectors(E, pairs):
    case_preseeded_vectors("cpu", pairs)


def case_preseeded_vectors(device, pairs):
    """online_update_tests.py:71-187: poke vectors_denoiser_cov_{u,v}, rebuild, time update."""
    from free_hunch_b200.covariance import CovarianceHessianBFGS
    g = np.load(H.ROOT + "/tests/golden/unit_tests.npz")
    p = f"t2_{pairs}_"
    d = 15
    bf = CovarianceHessianBFGS(init_denoiser_variance=1, init_noise_variance=80.0 ** 2, data_dim=d, device=device)
    bf.vectors_denoiser_cov_u = torch.from_numpy(g[p + "U"]).to(torch.complex128)
    bf.vectors_denoiser_cov_v = torch.from_numpy(g[p + "V"]).to(torch.complex128)
    bf.set_others_corresponding_to_current_denoiser_cov(80.0)
    cov, inv_cov, _, _ = bf.get_dense_matrices()
    assert rel_l2(cov, g[p + "pre_cov"]) < 1e-5
    assert rel_l2(inv_cov, g[p + "pre_inv_cov"]) < 1e-5
    U, V = bf.vectors_denoiser_cov_u, bf.vectors_denoiser_cov_v
    assert U.shape[1] == pairs and V.shape[1] == pairs
    x = torch.from_numpy(g[p + "x"])
    bf.update_time_step(x[None], 80.0, 79.0, (-x / 80.0 ** 2)[None])
    cov, inv_cov, _, _ = bf.get_dense_matrices()
    assert rel_l2(cov, g[p + "cov"]) < 1e-4
    assert rel_l2(inv_cov, g[p + "inv_cov"]) < 1e-4


def test_operators_and_cg_small(E, assets, monkeypatch):
    case_operators_and_cg_small(E, assets, "cpu", monkeypatch)


def case_operators_and_cg_small(E, assets, device, monkeypatch):
    """Operators (vs the oracle's FFT operators with a float64 OTF) and batched CG on 2 x 64 x 64 images, identity basis."""
    rs = np.random.RandomState(0)
    Hh, C = 64, 2
    ctx = E.Context(device, max_batch=2, channels=C, height=Hh, width=Hh)
    x = rs.standard_normal((2, C, Hh, Hh))
    for key in ("gaussian_ks61_std3", "motion_ks61_std05"):
        op = E.Operator.conv(ctx, assets[key], 1)
        oo = fo.GaussianBlurOracle(assets[key], 0.1, image_size=Hh, otf_dtype=np.complex128)
        assert rel_l2(op.apply(torch.tensor(x)), oo.A(x)) < 1e-6
        assert rel_l2(op.apply(torch.tensor(x), transpose=True), oo.AT(x)) < 1e-6
    if monkeypatch is not None:
        # the motion PSF through both pixel-space kernels: the tap list above (few planes), the column-run kernel here
        monkeypatch.setenv("FH_EMU_CONV_RUNS", "1")
        assert rel_l2(op.apply(torch.tensor(x)), oo.A(x)) < 1e-6
        assert rel_l2(op.apply(torch.tensor(x), transpose=True), oo.AT(x)) < 1e-6
        monkeypatch.delenv("FH_EMU_CONV_RUNS")
    op = E.Operator.conv(ctx, assets["bicubic_x4"], 4)
    oo = fo.SuperResolutionOracle(assets["bicubic_x4"], 0.1, 4, image_size=Hh, otf_dtype=np.complex128)
    ys = rs.standard_normal((2, C, Hh // 4, Hh // 4))
    assert rel_l2(op.apply(torch.tensor(x)), oo.A(x)) < 1e-6
    assert rel_l2(op.apply(torch.tensor(ys), transpose=True), oo.AT(ys)) < 1e-6
    # state with a low-rank part, identity basis, two different images
    var = np.exp(rs.uniform(np.log(1e-2), np.log(30.0), size=(C, Hh, Hh))).astype(np.float32)
    st = E.CovState(ctx, 2, E.BASIS_IDENTITY)
    st.init(torch.tensor(var), 400.0)
    ors = [fo.CovarianceOracle(var.reshape(-1), 400.0, var.size) for _ in range(2)]
    xs = rs.standard_normal((2, C, Hh, Hh)) * 20
    sc = -xs / 400.0 * 0.9
    mean, _ = st.time_update(torch.tensor(xs), torch.tensor(sc), 20.0, 8.0)
    for b in range(2):
        mo, _ = ors[b].update_time_step(xs[b:b + 1], 20.0, 8.0, sc[b:b + 1])
        assert rel_l2(mean[b], mo[0]) < 1e-5
    xn = xs + 0.3 * rs.standard_normal(xs.shape)
    den = lambda z: 0.5 * np.tanh(z) + 0.1 * z * var / (var + 64)      # monotone in z: dx.de > 0 as for a real denoiser
    m0, m1 = den(xs), den(xn)
    st.space_update(torch.tensor(m0), torch.tensor(m1), 8.0, torch.tensor(xs), torch.tensor(xn))
    for b in range(2):
        ors[b].update_space_step(m0[b:b + 1], m1[b:b + 1], 8.0, xs[b:b + 1], xn[b:b + 1])
    w = rs.standard_normal((2, C, Hh, Hh))
    sv = st.cov_dot(torch.tensor(w))
    for b in range(2):
        assert rel_l2(sv[b], ors[b].denoiser_cov_vector_dot(w[b:b + 1])[0]) < 1e-5
    # CG: fixed iterations and converged, SR operator, per-image convergence
    xc = np.clip(rs.standard_normal((2, C, Hh, Hh)), -1, 1)
    y = oo.A(xc) + 0.1 * rs.standard_normal(ys.shape)
    mu = 0.3 * xc
    sy2 = float(np.float32(0.1) * np.float32(0.1))
    for fixed, rtol in [(4, 1.0), (0, 1e-5)]:
        mat, info = E.cg_solve(st, op, torch.tensor(y), torch.tensor(mu), sy2, rtol, maxiter=200, fixed_iters=fixed)
        for b in range(2):
            def A_mm(u, b=b):
                u = u.reshape(y[b:b + 1].shape)
                return (sy2 * u + oo.A(ors[b].denoiser_cov_vector_dot(oo.AT(u)))).reshape(-1)
            u, io = fo.cg(A_mm, (y[b:b + 1] - oo.A(mu[b:b + 1])).reshape(-1), rtol=rtol, maxiter=200, fixed_iters=fixed or None)
            assert abs(info["niter"][b] - io["niter"]) <= 1
            assert info["optimal"][b] == io["optimal"]
            assert rel_l2(mat[b], oo.AT(u.reshape(y[b:b + 1].shape))[0]) < 1e-5


def case_project_to_diagonal(E, device, dct, assets=None):
    """update_space_step with project_to_diagonal (online_update_bfgs.py:274-277) on a state that already has a low-rank
    part: Sigma.v afterwards, and after a further time update, against the oracle (which follows the reference's Sigma side
    line by line).  Identity basis at 2 x 32 x 32 or the DCT basis at 3 x 256 x 256."""
    rs = np.random.RandomState(3)
    if dct:
        var = assets["dct_variance"]
        C, Hh = 3, 256
        ora = fo.CovarianceOracleDCT(var, 400.0, var.size)
        basis = E.BASIS_DCT
    else:
        C, Hh = 2, 32
        var = np.exp(rs.uniform(np.log(1e-2), np.log(30.0), size=(C, Hh, Hh))).astype(np.float32)
        ora = fo.CovarianceOracle(var.reshape(-1), 400.0, var.size)
        basis = E.BASIS_IDENTITY
    shape = (1, C, Hh, Hh)
    ctx = E.Context(device, max_batch=1, channels=C, height=Hh, width=Hh)
    st = E.CovState(ctx, 1, basis)
    st.init(torch.tensor(var), 400.0)
    den = lambda z: 0.5 * np.tanh(z) + 0.1 * z
    x = rs.standard_normal(shape) * 8
    st.time_update(None, None, 20.0, 8.0, only_covariance=True)
    ora.update_time_step(np.zeros(shape), 20.0, 8.0, np.zeros(shape), only_covariance=True)
    for step in range(3):
        xn = x + 0.3 * rs.standard_normal(shape)
        m0, m1 = den(x), den(xn)
        project = step == 2
        ora.project_to_diagonal = project
        st.space_update(torch.tensor(m0), torch.tensor(m1), 8.0, torch.tensor(x), torch.tensor(xn), project_to_diagonal=project)
        ora.update_space_step(m0, m1, 8.0, x, xn)
        x = xn
    assert st.rank == 4                                  # the projected pair did not add rows
    w = rs.standard_normal(shape)
    e1 = rel_l2(st.cov_dot(torch.tensor(w).float()), ora.denoiser_cov_vector_dot(w))
    st.time_update(None, None, 8.0, 5.0, only_covariance=True)
    ora.update_time_step(np.zeros(shape), 8.0, 5.0, np.zeros(shape), only_covariance=True)
    e2 = rel_l2(st.cov_dot(torch.tensor(w).float()), ora.denoiser_cov_vector_dot(w))
    assert e1 < 1e-4 and e2 < 1e-4, (e1, e2)
    return e1, e2


def case_drop_vectors(device, dct, assets=None, max_vector_count=2, steps=4):
    """drop_vectors(0 < n < k) (online_update_bfgs.py:233-245, called from :309-310) through the product classes: the
    reference keeps the last n columns of ITS factors; the product follows them with ref_factors.ReferenceFactorTracker
    and moves the state to their basis.  Sigma.v after every update against the oracle (which restates the reference's
    factor algebra line by line), and the rank stays bounded.  Identity basis at 2 x 16 x 16 or DCT at 3 x 256 x 256."""
    from free_hunch_b200.covariance import CovarianceHessianBFGS, CovarianceHessianBFGSDCT
    rs = np.random.RandomState(3)
    if dct:
        var = assets["dct_variance"]
        C, Hh = 3, 256
        ora = fo.CovarianceOracleDCT(var, 400.0, var.size, max_vector_count=max_vector_count)
        cov = CovarianceHessianBFGSDCT(None, 400.0, var.size, max_vector_count=max_vector_count,
                                       dct_variance=torch.tensor(var), device=device)
    else:
        C, Hh = 2, 16
        var = np.exp(rs.uniform(np.log(1e-2), np.log(30.0), size=(C, Hh, Hh))).astype(np.float32)
        ora = fo.CovarianceOracle(var.reshape(-1), 400.0, var.size, max_vector_count=max_vector_count)
        cov = CovarianceHessianBFGS(var.reshape(-1), 400.0, var.size, max_vector_count=max_vector_count, device=device,
                                    image_shape=(C, Hh, Hh))
    shape = (1, C, Hh, Hh)
    den = lambda z: 0.5 * np.tanh(z) + 0.1 * z
    sigs = [20.0, 8.0, 6.0, 5.0, 4.0, 3.0]
    x = rs.standard_normal(shape) * 8
    zero = np.zeros(shape)
    errs = []
    for i in range(steps):
        s, sn = sigs[i], sigs[i + 1]
        ora.update_time_step(zero, s, sn, zero, only_covariance=True)
        cov.update_time_step(torch.tensor(zero), s, sn, torch.tensor(zero), only_covariance=True)
        xn = x + 0.3 * rs.standard_normal(shape)
        m0, m1 = den(x), den(xn)
        ora.update_space_step(m0, m1, sn, x, xn)
        cov.update_space_step(torch.tensor(m0), torch.tensor(m1), sn, torch.tensor(x), torch.tensor(xn))
        assert cov.vector_count == ora.cov_u.shape[1] == min(i + 1, max_vector_count)
        w = rs.standard_normal(shape)
        errs.append(rel_l2(cov.denoiser_cov_vector_dot(torch.tensor(w).float()), ora.denoiser_cov_vector_dot(w)))
        x = xn
    assert max(errs) < 1e-4, errs
    return errs


def test_drop_vectors_identity_basis(E):
    case_drop_vectors("cpu", False)


def test_drop_vectors_dct_basis(E, assets):
    case_drop_vectors("cpu", True, assets, steps=3)


def test_drop_vectors_needs_the_factor_history(E):
    """Without max_vector_count at construction there is no factor history to truncate: loud refusal, not a guess."""
    from free_hunch_b200.covariance import CovarianceHessianBFGS
    cov = CovarianceHessianBFGS(np.ones(64, np.float32), 400.0, 64, device="cpu", image_shape=(1, 8, 8))
    rs = np.random.RandomState(0)
    x = rs.standard_normal((1, 1, 8, 8))
    for _ in range(2):
        xn = x + 0.3 * rs.standard_normal(x.shape)
        cov.update_space_step(torch.tensor(0.5 * np.tanh(x)), torch.tensor(0.5 * np.tanh(xn)), 20.0, torch.tensor(x), torch.tensor(xn))
        x = xn
    with pytest.raises(NotImplementedError, match="factor history"):
        cov.drop_vectors(1, 20.0)
    cov.drop_vectors(0, 20.0)
    assert cov.vector_count == 0


def test_project_to_diagonal_identity_basis(E):
    case_project_to_diagonal(E, "cpu", False)


def test_argument_errors(E):
    from free_hunch_b200 import _lib
    ctx = E.Context("cpu", max_batch=1, channels=1, height=8, width=8)
    with pytest.raises(_lib.FreeHunchLibraryError, match="256x256"):
        E.CovState(ctx, 1, E.BASIS_DCT)
    with pytest.raises(_lib.FreeHunchLibraryError):
        E.Operator.conv(ctx, np.zeros((4, 4), np.float32), 1)       # even kernel size
    with pytest.raises(_lib.FreeHunchLibraryError):
        E.CovState(ctx, 2, E.BASIS_IDENTITY)                         # batch > max_batch
    with pytest.raises(_lib.FreeHunchLibraryError):
        ctx.set_cg_mode(7)                                           # not one of FH_CG_AUTO / STREAMING / RESIDENT
    # the asynchronous solve exists only where the resident kernel applies: identity basis -> FH_ESTATE -> None
    st = E.CovState(ctx, 1, E.BASIS_IDENTITY)
    st.init(torch.ones(64), 4.0)
    op = E.Operator.inpaint(ctx, np.ones((8, 8), np.float32))
    y = torch.zeros(1, 1, 8, 8)
    assert E.cg_solve_async(st, op, y, y, 0.01, 1.0) is None
    # a destroyed state's buffers are reused by the next state of the same batch (no cudaMalloc / cudaFree per image)
    st.close()
    st2 = E.CovState(ctx, 1, E.BASIS_IDENTITY)
    st2.init(torch.full((64,), 2.0), 9.0)
    assert st2.rank == 0 and rel_l2(st2.cov_dot(torch.ones(1, 1, 8, 8)).reshape(-1), 2.0 * np.ones(64)) < 1e-6


def test_fused_dct_basis_path(E, assets, monkeypatch):
    """DCT basis at 1 x 256 x 256 (one channel keeps the emulation affordable): basis-layout state, fused
    Sigma.v, time/space update, and the fused CG iteration (mask folded in, separable blur folded in, generic operator) vs the oracle.
    Motion blur shares the generic path with super-resolution; the residual-test stop is covered by the -m gpu tests."""
    monkeypatch.setenv("FH_EMU_ROWS", "8")
    rs = np.random.RandomState(3)
    var = assets["dct_variance"][:1]                           # (1,256,256)
    shape = (1, 1, 256, 256)
    s0, s1 = 80.2878, 8.463
    ctx = E.Context("cpu", max_batch=1, channels=1, height=256, width=256)
    st = E.CovState(ctx, 1, E.BASIS_DCT)
    st.init(torch.tensor(var), s0 ** 2)
    ora = fo.CovarianceOracleDCT(var, s0 ** 2, var.size)
    den = lambda x, s: fo.idct2(var.astype(np.float64) / (var + s ** 2) * fo.dct2(x)) + 0.02 * np.tanh(x)
    w = rs.standard_normal(shape)
    assert rel_l2(st.cov_dot(torch.tensor(w)), ora.denoiser_cov_vector_dot(w)) < 1e-5          # k = 0
    x = rs.standard_normal(shape) * s0
    mu = den(x, s0)
    score = (mu - x) / s0 ** 2
    m_gpu, sc_gpu = st.time_update(torch.tensor(x), torch.tensor(score), s0, s1)
    m_cpu, sc_cpu = ora.update_time_step(x, s0, s1, score)
    assert rel_l2(m_gpu, m_cpu) < 1e-4 and rel_l2(sc_gpu, sc_cpu) < 1e-5
    xn = x + (s1 - s0) * (x - mu) / s0
    mun = den(xn, s1)
    st.space_update(torch.tensor(m_cpu), torch.tensor(mun), s1, torch.tensor(x), torch.tensor(xn))
    ora.update_space_step(m_cpu, mun, s1, x, xn)
    assert st.rank == 2
    assert rel_l2(st.cov_dot(torch.tensor(w)), ora.denoiser_cov_vector_dot(w)) < 1e-5          # k = 1
    # exported vectors come back in the natural (ki, kj) order
    pi, Z, Gz, Gs, delta = st.export()
    d_ref = 1.0 / (1.0 / var.astype(np.float64) - 1.0 / s0 ** 2 + 1.0 / s1 ** 2)
    assert rel_l2(delta.reshape(var.shape), d_ref) < 1e-6
    assert rel_l2(Z[0, 0].reshape(256, 256), fo.dct2(xn - x)[0, 0]) < 1e-5
    xc = np.clip(rs.standard_normal(shape), -1, 1)
    mu0 = 0.3 * xc
    sy2 = float(np.float32(0.1) * np.float32(0.1))
    mask = (rs.uniform(size=(256, 256)) > 0.7).astype(np.float32)
    cases = [
        ("inpainting", E.Operator.inpaint(ctx, mask), fo.InpaintingOracle(np.broadcast_to(mask, shape), 0.1)),
        ("gaussian", E.Operator.conv(ctx, assets["gaussian_ks61_std3"], 1), fo.GaussianBlurOracle(assets["gaussian_ks61_std3"], 0.1, otf_dtype=np.complex128)),
        ("sr", E.Operator.conv(ctx, assets["bicubic_x4"], 4), fo.SuperResolutionOracle(assets["bicubic_x4"], 0.1, 4, otf_dtype=np.complex128)),
    ]
    for name, op, oo in cases:
        if name == "inpainting":
            y = (xc + 0.1 * rs.standard_normal(shape)) * mask
        else:
            ay = oo.A(xc)
            y = ay + 0.1 * rs.standard_normal(ay.shape)
        mat, info = E.cg_solve(st, op, torch.tensor(y), torch.tensor(mu0), sy2, 1.0, maxiter=50, fixed_iters=2)
        mat_cpu, io = fo.mat_solver(oo, y, mu0, ora, s1, fixed_iters=2)
        assert info["niter"][0] == 2, name
        assert rel_l2(mat, mat_cpu) < 2e-5, (name, rel_l2(mat, mat_cpu))


def test_resident_cg_kernel(E, assets):
    """The one-launch resident CG solve (fh_cg_resident.cuh) at 3 x 256 x 256, all 128 CTAs of its cooperative grid
    emulated concurrently (32 768 fibers, tools/cuda_emu.h): 6-row strips that straddle channel planes, Z slices resident in shared memory, grid
    barriers, in-kernel initial residual / step / stop logic -- vs the oracle and vs the five-launch streaming path."""
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
    m_gpu, _ = st.time_update(torch.tensor(x), torch.tensor(score), s0, s1)
    m_cpu, _ = ora.update_time_step(x, s0, s1, score)
    xn = x + (s1 - s0) * (x - mu) / s0
    mun = den(xn, s1)[0] + 0.02 * np.tanh(xn)
    st.space_update(torch.tensor(m_cpu), torch.tensor(mun), s1, torch.tensor(x), torch.tensor(xn))
    ora.update_space_step(m_cpu, mun, s1, x, xn)
    assert st.rank == 2
    xc = np.clip(rs.standard_normal(shape), -1, 1)
    mu0 = 0.3 * xc
    sy2 = float(np.float32(0.1) * np.float32(0.1))
    mask = (rs.uniform(size=(256, 256)) > 0.7).astype(np.float32)
    cases = [
        ("inpainting", E.Operator.inpaint(ctx, mask), fo.InpaintingOracle(np.broadcast_to(mask, shape), 0.1)),
        ("gaussian", E.Operator.conv(ctx, assets["gaussian_ks61_std3"], 1), fo.GaussianBlurOracle(assets["gaussian_ks61_std3"], 0.1, otf_dtype=np.complex128)),
    ]
    for name, op, oo in cases:
        if name == "inpainting":
            y = (xc + 0.1 * rs.standard_normal(shape)) * mask
        else:
            ay = oo.A(xc)
            y = ay + 0.1 * rs.standard_normal(ay.shape)
        ctx.set_cg_mode(E.CG_RESIDENT)
        mat, info = E.cg_solve(st, op, torch.tensor(y), torch.tensor(mu0), sy2, 1.0, maxiter=50, fixed_iters=2)
        assert ctx.cg_last_mode == E.CG_RESIDENT
        mat_cpu, io = fo.mat_solver(oo, y, mu0, ora, s1, fixed_iters=2)
        assert info["niter"][0] == 2, name
        assert rel_l2(mat, mat_cpu) < 2e-5, (name, rel_l2(mat, mat_cpu))
        if name == "inpainting":
            ctx.set_cg_mode(E.CG_STREAMING)
            mat2, info2 = E.cg_solve(st, op, torch.tensor(y), torch.tensor(mu0), sy2, 1.0, maxiter=50, fixed_iters=2)
            assert ctx.cg_last_mode == E.CG_STREAMING
            assert rel_l2(mat, mat2) < 1e-5
            assert abs(info["residual_norm"][0] - info2["residual_norm"][0]) <= 1e-4 * info2["residual_norm"][0]
    ctx.set_cg_mode(E.CG_AUTO)


def test_resident_cg_kernel_rank_above_32(E, assets):
    """fh_k_cg_resident<true>: rank 34 (17 BFGS pairs) -- 32 rows of Z resident, two re-read from global memory, 64-row
    coefficient stage and owner mapping -- one CG iteration against the oracle, Gaussian blur folded in."""
    rs = np.random.RandomState(9)
    var = assets["dct_variance"]
    shape = (1,) + var.shape
    s0, s1 = 80.2878, 6.5
    ctx = E.Context("cpu", max_batch=1)
    st = E.CovState(ctx, 1, E.BASIS_DCT)
    st.init(torch.tensor(var), s0 ** 2)
    ora = fo.CovarianceOracleDCT(var, s0 ** 2, var.size)
    st.time_update(None, None, s0, s1, only_covariance=True)
    ora.update_time_step(np.zeros(shape), s0, s1, np.zeros(shape), only_covariance=True)
    den = lambda x: fo.idct2(var.astype(np.float64) / (var + s1 ** 2) * fo.dct2(x)) + 0.02 * np.tanh(x)
    x = np.clip(rs.standard_normal(shape), -1, 1) + s1 * rs.standard_normal(shape)
    for _ in range(17):
        xn = x + 0.3 * s1 * rs.standard_normal(shape)
        m0, m1 = den(x), den(xn)
        st.space_update(torch.tensor(m0), torch.tensor(m1), s1, torch.tensor(x), torch.tensor(xn))
        ora.update_space_step(m0, m1, s1, x, xn)
        x = xn
    assert st.rank == 34
    op = E.Operator.conv(ctx, assets["gaussian_ks61_std3"], 1)
    oo = fo.GaussianBlurOracle(assets["gaussian_ks61_std3"], 0.1, otf_dtype=np.complex128)
    xc = np.clip(rs.standard_normal(shape), -1, 1)
    y = oo.A(xc) + 0.1 * rs.standard_normal(shape)
    mu0 = 0.3 * xc
    sy2 = float(np.float32(0.1) * np.float32(0.1))
    ctx.set_cg_mode(E.CG_RESIDENT)
    mat, info = E.cg_solve(st, op, torch.tensor(y), torch.tensor(mu0), sy2, 1.0, maxiter=50, fixed_iters=1)
    assert ctx.cg_last_mode == E.CG_RESIDENT and info["niter"][0] == 1
    mat_cpu, _ = fo.mat_solver(oo, y, mu0, ora, s1, fixed_iters=1)
    assert rel_l2(mat, mat_cpu) < 2e-5, rel_l2(mat, mat_cpu)
    ctx.set_cg_mode(E.CG_AUTO)
