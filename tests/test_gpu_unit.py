"""The reference's own unit tests (conditioning_utils/online_update_tests.py:3-374, F3-adapted: leading batch dimension,
complex cast of the pre-seeded vectors) and the identity-basis code path (CovarianceHessianBFGS without the DCT: plain
low-rank kernels fh_k_dots / fh_k_lowrank_apply, fh_state_set_cov_vectors, fh_k_core_from_cov, fh_state_export, the
generic three-kernel CG fh_k_cg_q / update1 / update2) ON HARDWARE through the C ABI -- in round 1 these ran only under
the CPU emulation.  Goldens: tests/golden/unit_tests.npz, produced by the real reference.

Plus BASELINE configs[3]'s "dense-matrix vs memory-efficient update" leg: the reference's dense forms
(online_update_bfgs.py:377-463, oracle/fh_oracle.py::dense_update_*) hold d x d matrices -- 154 GB at d = 196 608 (SURVEY
F12) -- so the comparison runs at 3 x 32 x 32 (d = 3072, 75 MB per float64 matrix): same Sigma.v, same time-update mean,
and the time of one update on each side."""
import time

import numpy as np
import pytest
import torch

import helpers as H
from helpers import rel_l2, record

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not H.HAS_CUDA, reason="needs a CUDA device")]

from oracle import fh_oracle as fo  # noqa: E402
import test_kernels_emu as cases  # noqa: E402   (the case bodies are shared with the CPU emulation run)


@pytest.fixture(scope="module")
def E():
    from free_hunch_b200 import engine, _lib
    assert _lib.lib()._name.endswith("libfreehunch_b200.so"), "GPU tests must run on the real sm_100a library"
    return engine


def test_reference_unit_test_time_and_space_identity_basis(E):
    cases.case_time_and_space_identity_basis("cuda")


@pytest.mark.parametrize("pairs", [1, 2, 4, 8])
def test_reference_unit_test_preseeded_vectors(E, pairs):
    cases.case_preseeded_vectors("cuda", pairs)


def test_identity_basis_operators_and_generic_cg(E, assets):
    cases.case_operators_and_cg_small(E, assets, "cuda", None)


@pytest.mark.parametrize("dct", [False, True])
def test_project_to_diagonal(E, assets, dct):
    e1, e2 = cases.case_project_to_diagonal(E, "cuda", dct, assets)
    record("project_to_diagonal", dct=dct, sigv_after_update=e1, sigv_after_time_update=e2)


@pytest.mark.parametrize("dct", [False, True])
def test_drop_vectors_keeps_the_reference_columns(E, assets, dct):
    """drop_vectors(0 < n < k), online_update_bfgs.py:233-245: max_vector_count = 2 over 5 (identity) / 4 (DCT) space
    updates, Sigma.v against the oracle's truncated factor model after every update."""
    errs = cases.case_drop_vectors("cuda", dct, assets, steps=4 if dct else 5)
    record("drop_vectors", dct=dct, sigv_errs=[float(e) for e in errs])


def test_dense_vs_memory_efficient_d3072(E):
    """configs[3], dense leg: 6 alternating time + space updates at d = 3 x 32 x 32 on both representations."""
    rs = np.random.RandomState(7)
    C, Hh = 3, 32
    d = C * Hh * Hh
    var = np.exp(rs.uniform(np.log(1e-2), np.log(30.0), size=d))
    sig = [20.0, 14.0, 9.0, 6.0, 4.0, 2.5, 1.5]
    ctx = E.Context("cuda", max_batch=1, channels=C, height=Hh, width=Hh)
    st = E.CovState(ctx, 1, E.BASIS_IDENTITY)
    st.init(torch.from_numpy(var), sig[0] ** 2)
    cov = np.diag(var)[None]
    inv_cov = np.diag(1.0 / var)[None]
    hes = (cov / sig[0] ** 2 - np.eye(d)) / sig[0] ** 2
    inv_hes = np.linalg.inv(hes)
    den = lambda z, s: z * var / (var + s * s) + 0.05 * np.tanh(z)       # monotone: dx.de > 0 like a real denoiser
    x = rs.standard_normal((1, d)) * sig[0]
    t_gpu, t_cpu, errs = [], [], []
    for i in range(6):
        s, sn = sig[i], sig[i + 1]
        mu = den(x, s)
        score = (mu - x) / s ** 2
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mean_g, _ = st.time_update(torch.from_numpy(x).reshape(1, C, Hh, Hh), torch.from_numpy(score).reshape(1, C, Hh, Hh), s, sn)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        cov, inv_cov, hes, inv_hes, _, mean_d = fo.dense_update_covariance(x, cov, inv_cov, hes, inv_hes, score, s, sn)
        t2 = time.perf_counter()
        xn = x + (sn - s) * (x - mu) / s + 0.05 * sn * rs.standard_normal((1, d))
        mun = den(xn, sn)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        st.space_update(mean_g, torch.from_numpy(mun).reshape(1, C, Hh, Hh), sn, torch.from_numpy(x).reshape(1, C, Hh, Hh), torch.from_numpy(xn).reshape(1, C, Hh, Hh))
        torch.cuda.synchronize(); t4 = time.perf_counter()
        cov, inv_cov, hes, inv_hes = fo.dense_update_bfgs(cov, inv_cov, mean_d, mun, sn, xn - x)
        t5 = time.perf_counter()
        w = rs.standard_normal((1, d))
        e_mean = rel_l2(mean_g.reshape(1, d), mean_d)
        e_sv = rel_l2(st.cov_dot(torch.from_numpy(w).reshape(1, C, Hh, Hh)).reshape(1, d), (cov[0] @ w[0])[None])
        errs.append((e_mean, e_sv))
        t_gpu.append((t1 - t0) + (t4 - t3)); t_cpu.append((t2 - t1) + (t5 - t4))
        x = xn
    record("dense_vs_lowrank_d3072", errs=[[float(a), float(b)] for a, b in errs], rank=st.rank,
           gpu_ms_per_update_pair=1e3 * float(np.median(t_gpu)), dense_cpu_ms_per_update_pair=1e3 * float(np.median(t_cpu)),
           dense_bytes_per_matrix=d * d * 8, lowrank_state_bytes=(2 + st.rank) * d * 4)
    assert st.rank == 12
    assert max(e for e, _ in errs) < 1e-4 and max(e for _, e in errs) < 1e-4, errs
