"""CPU oracle for the Free Hunch guided-sampling hot path (TEST INFRASTRUCTURE ONLY).

This module is a numpy/scipy restatement, in the reference's own precision (complex128
state, float64 CG), of the algorithm that AaltoML/free-hunch runs under
``conditioning_mechanism=online_covariance``.  It exists to *check* the CUDA product in
``free_hunch_b200`` -- it is never on the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference from
``/root/reference`` (with the shims of SURVEY.md Appendix B), runs it on seeded inputs and
stores the outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays those
inputs through this file.  The one third-party boundary that cannot be pinned against the
reference's own dependency is ``torch-dct==0.1.6`` (not installed anywhere here): the
orthonormal DCT-II/III it computes is uniquely defined and is restated with
``scipy.fft.dctn(norm='ortho')``, the transform the reference itself uses for the same
purpose at ``conditioning_utils/utils.py:94,101``.

Every function cites the reference file:line (relative to the reference root) it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np
import scipy.fft as sfft
from scipy.linalg import sqrtm as _scipy_sqrtm

C128 = np.complex128


# ----------------------------------------------------------------------------------------
# Orthonormal 2-D DCT over the last two axes (torch_dct.dct_2d / idct_2d, norm='ortho')
# call sites: conditioning_utils/online_update_bfgs.py:352,355,358,360,363-366,373-374
# ----------------------------------------------------------------------------------------
def dct2(x: np.ndarray) -> np.ndarray:
    return sfft.dctn(np.asarray(x), type=2, norm="ortho", axes=(-2, -1))


def idct2(x: np.ndarray) -> np.ndarray:
    return sfft.idctn(np.asarray(x), type=2, norm="ortho", axes=(-2, -1))


# ----------------------------------------------------------------------------------------
# CovarianceHessianBFGS  (conditioning_utils/online_update_bfgs.py:7-336)
# ----------------------------------------------------------------------------------------
def _sqrtm(a: np.ndarray) -> np.ndarray:
    """online_update_bfgs.py:67-71 -- scipy sqrtm, complex128, zeros for empty input."""
    if a.shape[0] != 0 and a.shape[1] != 0:
        return np.asarray(_scipy_sqrtm(a)).astype(C128)
    return np.zeros_like(a)


def woodbury_inverse_from_diag_plus_lowrank(diag_inv: np.ndarray, U: np.ndarray) -> np.ndarray:
    """(diag + U U^T)^-1 = diag_inv - V_inv V_inv^T; returns V_inv.  online_update_bfgs.py:87-104."""
    k = U.shape[-1]
    inner_inv = np.linalg.inv(np.eye(k) + U.T @ (diag_inv[:, None] * U)) if k else np.zeros((0, 0), C128)
    inner_inv = (inner_inv + inner_inv.T) / 2            # :98 (plain transpose)
    inner_inv_sqrt = _sqrtm(inner_inv)                   # :101
    return diag_inv[:, None] * (U @ inner_inv_sqrt)      # :103


def woodbury_inverse_from_diag_plus_lowrank_minus_lowrank(U, V, diag):
    """(diag + UU^T - VV^T)^-1 as diag_inv + U_inv U_inv^T - V_inv V_inv^T.  online_update_bfgs.py:106-119."""
    diag_inv = 1 / diag
    V_inv = woodbury_inverse_from_diag_plus_lowrank(diag_inv, U)
    k = V.shape[-1]
    K = V_inv.T @ V
    if k:
        inner_inv = np.linalg.inv(np.eye(k) - V.T @ (diag_inv[:, None] * V) + K.T @ K)   # :115 (not symmetrised)
    else:
        inner_inv = np.zeros((0, 0), C128)
    inner_inv_sqrt = _sqrtm(inner_inv)
    V_inner = V @ inner_inv_sqrt
    U_inv = diag_inv[:, None] * V_inner - V_inv @ (V_inv.T @ V_inner)                    # :118
    return diag_inv, U_inv, V_inv


class CovarianceOracle:
    """Identity-basis covariance/Hessian state: four families {cov, inv_cov, hes, inv_hes}, each
    diag + U U^T - V V^T with plain transposes, complex128.  online_update_bfgs.py:7-336."""

    def __init__(self, init_denoiser_variance, init_noise_variance, data_dim,
                 max_vector_count=None, init_denoiser_cov_u=None, project_to_diagonal=False):
        d = data_dim
        z = lambda: np.zeros((d, 0), C128)
        self.cov_u = z() if init_denoiser_cov_u is None else np.asarray(init_denoiser_cov_u, C128)   # :16-18
        self.cov_v = z()
        self.diag_cov = np.ones(d, C128) * np.asarray(init_denoiser_variance).reshape(-1)            # :27
        self.diag_inv_cov = 1 / self.diag_cov                                                        # :28
        idv = np.asarray(init_denoiser_variance).reshape(-1)
        if idv.dtype == np.float32:
            # :29 is evaluated on the float32 `dct_variance` tensor BEFORE the complex128 multiply, so the
            # reference's initial Hessian diagonal is float32-rounded (visible at the 1e-4 level in the first
            # time-updated means).  Reproduced on purpose.
            nv = np.float32(init_noise_variance)
            h0 = ((idv / nv - np.float32(1)) / nv).astype(np.float32)
        else:
            h0 = (idv / init_noise_variance - 1) / init_noise_variance
        self.diag_hes = h0 * np.ones(d, C128)                                                        # :29
        self.diag_inv_hes = 1 / self.diag_hes                                                        # :30
        self.inv_cov_u, self.inv_cov_v = z(), z()
        self.hes_u, self.hes_v, self.inv_hes_u, self.inv_hes_v = z(), z(), z(), z()
        self.data_dim = d
        self.max_vector_count = max_vector_count
        self.project_to_diagonal = project_to_diagonal
        self.set_others_corresponding_to_current_denoiser_cov(np.sqrt(init_noise_variance))          # :37

    # -- basis hooks (identity here; DCT subclass overrides) :314-318
    def transform(self, x):
        return x

    def inverse_transform(self, x):
        return x

    # -- mat-vecs :194-231
    @staticmethod
    def _mv(diag, U, V, v):
        return diag * v + U @ (U.T @ v) - V @ (V.T @ v)

    def _denoiser_cov_vector_dot(self, v):
        return self._mv(self.diag_cov, self.cov_u, self.cov_v, v)

    def denoiser_cov_vector_dot(self, v):
        v = np.asarray(v)
        return self._denoiser_cov_vector_dot(v.astype(C128).reshape(-1)).real.reshape(v.shape).astype(v.dtype)

    def inv_denoiser_cov_vector_dot(self, v):
        v = np.asarray(v)
        return self._mv(self.diag_inv_cov, self.inv_cov_u, self.inv_cov_v, v.astype(C128).reshape(-1)).real.reshape(v.shape)

    def hessian_vector_dot(self, v):
        v = np.asarray(v)
        return self._mv(self.diag_hes, self.hes_u, self.hes_v, v.astype(C128).reshape(-1)).real.reshape(v.shape)

    def inv_hessian_vector_dot(self, v):
        v = np.asarray(v)
        return self._mv(self.diag_inv_hes, self.inv_hes_u, self.inv_hes_v, v.astype(C128).reshape(-1)).real.reshape(v.shape)

    # -- time update :153-192
    def update_time_step(self, x_t, sigma_t, sigma_tnext, score_t, only_covariance=False):
        mean, score = self._update_time_step(x_t, sigma_t, sigma_tnext, score_t, only_covariance)
        return mean.real, score.real

    def _update_time_step(self, x_t, sigma_t, sigma_tnext, score_t, only_covariance=False):
        shape = np.shape(x_t)
        assert shape[0] == 1, "Batch size must be 1"                                   # :161
        x = np.asarray(x_t, C128).reshape(-1)
        s = np.asarray(score_t, C128).reshape(-1)
        # :166/:172 multiply the python-float shift by torch.ones(d) (float32): the shift is float32-rounded
        self.diag_inv_cov = self.diag_inv_cov + float(np.float32(sigma_tnext ** (-2) - sigma_t ** (-2)))   # :166
        self.diag_cov, self.cov_u, self.cov_v = woodbury_inverse_from_diag_plus_lowrank_minus_lowrank(
            self.inv_cov_u, self.inv_cov_v, self.diag_inv_cov)                                       # :167
        if not only_covariance:
            new_diag_inv_hes = self.diag_inv_hes - float(np.float32(sigma_tnext ** 2 - sigma_t ** 2))   # :172
            new_diag_hes, new_u, new_v = woodbury_inverse_from_diag_plus_lowrank_minus_lowrank(
                self.inv_hes_u, self.inv_hes_v, new_diag_inv_hes)                                   # :173
            old = self._mv(self.diag_inv_hes, self.inv_hes_u, self.inv_hes_v, s)                    # :176
            new_score = self._mv(new_diag_hes, new_u, new_v, old).real + 0j                         # :177
            new_mean = (x + sigma_tnext ** 2 * new_score).real + 0j                                 # :179
            self.diag_inv_hes = new_diag_inv_hes
            self.diag_hes, self.hes_u, self.hes_v = new_diag_hes, new_u, new_v
            return new_mean.reshape(shape), new_score.reshape(shape)
        return (x + 0j).reshape(shape), (x + 0j).reshape(shape)                                     # :187-188

    # -- space (BFGS) update :247-312
    def update_space_step(self, mean_at_x, mean_at_xnext, sigma_t, x, xnext):
        shape = np.shape(x)
        assert shape[0] == 1, "Batch size must be 1"                                   # :255
        x = np.asarray(x, C128).reshape(-1)
        xnext = np.asarray(xnext, C128).reshape(-1)
        m0 = np.asarray(mean_at_x, C128).reshape(-1)
        m1 = np.asarray(mean_at_xnext, C128).reshape(-1)
        dx = xnext - x
        de = sigma_t ** 2 * (m1 - m0)
        gamma = 1 / (dx @ de)                                                            # :263
        cdx = self._denoiser_cov_vector_dot(dx)                                          # :268
        dcd = cdx @ dx                                                                   # :270
        v = cdx / np.sqrt(dcd + 0j)                                                      # :272 (complex sqrt)
        u = de * np.sqrt(gamma + 0j)                                                     # :273
        if self.project_to_diagonal:                                                     # :274-277
            new_diag = self.diag_cov + u * u - v * v
            new_u, new_v = self.cov_u, self.cov_v
        else:                                                                            # :279-281
            new_diag = self.diag_cov
            new_u = np.concatenate([self.cov_u, u[:, None]], axis=-1)
            new_v = np.concatenate([self.cov_v, v[:, None]], axis=-1)
        inv = woodbury_inverse_from_diag_plus_lowrank_minus_lowrank(new_u, new_v, new_diag)          # :289
        new_diag_hes = (new_diag / sigma_t ** 2 - 1) / sigma_t ** 2                                   # :294
        new_hes_u = np.concatenate([self.hes_u, (u / sigma_t ** 2)[:, None]], axis=-1)                # :295-298
        new_hes_v = np.concatenate([self.hes_v, (v / sigma_t ** 2)[:, None]], axis=-1)
        inv_h = woodbury_inverse_from_diag_plus_lowrank_minus_lowrank(new_hes_u, new_hes_v, new_diag_hes)  # :300
        self.diag_cov, self.cov_u, self.cov_v = new_diag, new_u, new_v
        self.diag_inv_cov, self.inv_cov_u, self.inv_cov_v = inv
        self.diag_hes, self.hes_u, self.hes_v = new_diag_hes, new_hes_u, new_hes_v
        self.diag_inv_hes, self.inv_hes_u, self.inv_hes_v = inv_h
        if self.max_vector_count is not None:
            self.drop_vectors(self.max_vector_count, sigma_t)                            # :309-310

    def drop_vectors(self, max_vector_count, sigma):                                     # :233-245
        if max_vector_count == 0:
            self.cov_u = np.zeros((self.data_dim, 0), C128)
            self.cov_v = np.zeros((self.data_dim, 0), C128)
            self.set_others_corresponding_to_current_denoiser_cov(sigma)
        elif self.cov_u.shape[-1] > max_vector_count:
            n = min(self.cov_u.shape[-1], max_vector_count)
            self.cov_u = self.cov_u[:, -n:]
            self.cov_v = self.cov_v[:, -n:]
            self.set_others_corresponding_to_current_denoiser_cov(sigma)

    def set_others_corresponding_to_current_denoiser_cov(self, sigma):                   # :327-330
        self.diag_inv_cov, self.inv_cov_u, self.inv_cov_v = \
            woodbury_inverse_from_diag_plus_lowrank_minus_lowrank(self.cov_u, self.cov_v, self.diag_cov)
        self.diag_hes = (self.diag_cov / sigma ** 2 - 1) / sigma ** 2
        self.hes_u, self.hes_v = self.cov_u / sigma ** 2, self.cov_v / sigma ** 2
        self.diag_inv_hes, self.inv_hes_u, self.inv_hes_v = \
            woodbury_inverse_from_diag_plus_lowrank_minus_lowrank(self.hes_u, self.hes_v, self.diag_hes)

    def get_dense_matrices(self):                                                        # :320-325
        def dense(diag, U, V):
            return np.diag(diag) + U @ U.T - V @ V.T
        return (dense(self.diag_cov, self.cov_u, self.cov_v),
                dense(self.diag_inv_cov, self.inv_cov_u, self.inv_cov_v),
                dense(self.diag_hes, self.hes_u, self.hes_v),
                dense(self.diag_inv_hes, self.inv_hes_u, self.inv_hes_v))


class CovarianceOracleDCT(CovarianceOracle):
    """Same state, all operations in the per-channel orthonormal 2-D DCT basis.
    online_update_bfgs.py:339-374.  ``dct_variance`` is the (C,H,W) array the reference loads from
    ``<data_dir>/dct_variance.pt`` (:343); ``use_precalculated_info=False`` replaces it with ones (:348)."""

    def __init__(self, dct_variance, init_noise_variance, data_dim, max_vector_count=None,
                 project_to_diagonal=False, use_precalculated_info=True):
        # float32 in both branches: torch.load gives a float32 tensor (:343), torch.ones is float32 (:348)
        var = np.asarray(dct_variance, np.float32) if use_precalculated_info else np.ones(data_dim, np.float32)
        self.dct_variance = var
        super().__init__(var.reshape(-1), init_noise_variance, data_dim,
                         max_vector_count=max_vector_count, project_to_diagonal=project_to_diagonal)

    def transform(self, x):
        return dct2(x)

    def inverse_transform(self, x):
        return idct2(x)

    def update_time_step(self, x_t, sigma_t, sigma_tnext, score_t, only_covariance=False):           # :357-360
        m, s = self._update_time_step(dct2(x_t), sigma_t, sigma_tnext, dct2(score_t), only_covariance)
        return idct2(m.real), idct2(s.real)

    def update_space_step(self, mean_at_x, mean_at_xnext, sigma_t, x, xnext):                        # :362-368
        super().update_space_step(dct2(mean_at_x), dct2(mean_at_xnext), sigma_t, dct2(x), dct2(xnext))

    def denoiser_cov_vector_dot(self, v):                                                            # :370-374
        v = np.asarray(v)
        vh = dct2(v).astype(C128).reshape(-1)
        return idct2(self._denoiser_cov_vector_dot(vh).real.reshape(v.shape).astype(v.dtype))


# ----------------------------------------------------------------------------------------
# Dense forms  (online_update_bfgs.py:377-463; notebook twins
# notebooks/diffusion_gaussian_mixture_utils.py:839-923)
# ----------------------------------------------------------------------------------------
def dense_update_covariance(samples, cov, inv_cov, hes, inv_hes, score, sigma_t, sigma_tnext):
    """Batched (bs,d,d) time update.  online_update_bfgs.py:398-412 with schedule(t)=sigma."""
    d = samples.shape[-1]
    eye = np.eye(d)
    new_inv_cov = inv_cov + (sigma_tnext ** (-2) - sigma_t ** (-2)) * eye
    new_cov = np.linalg.inv(new_inv_cov)
    new_inv_hes = inv_hes - (sigma_tnext ** 2 - sigma_t ** 2) * eye
    new_hes = np.linalg.inv(new_inv_hes)
    new_score = (new_hes @ inv_hes @ score[..., None])[..., 0]
    new_mean = samples + sigma_tnext ** 2 * new_score
    return new_cov, new_inv_cov, new_hes, new_inv_hes, new_score, new_mean


def dense_update_bfgs(cov, inv_cov, mean_at_x, mean_at_xnext, sigma_t, dx):
    """Batched (bs,d,d) BFGS space update.  online_update_bfgs.py:451-463."""
    d = dx.shape[-1]
    eye = np.eye(d)
    de = sigma_t ** 2 * (mean_at_xnext - mean_at_x)
    dxc, dxr = dx[..., :, None], dx[..., None, :]
    dec, der = de[..., :, None], de[..., None, :]
    gamma = 1 / (dxr @ dec)
    new_cov = cov - cov @ dxc @ dxr @ cov / (dxr @ cov @ dxc) + dec @ der * gamma
    new_inv_cov = (eye - dxc @ der * gamma) @ inv_cov @ (eye - dec @ dxr * gamma) + dxc @ dxr * gamma
    new_hes = (new_cov / sigma_t ** 2 - eye) / sigma_t ** 2
    new_inv_hes = np.linalg.inv(new_hes + 1e-10 * eye)
    return new_cov, new_inv_cov, new_hes, new_inv_hes


# ----------------------------------------------------------------------------------------
# CG  (conditioning_utils/cg.py:118-292, live body :230-292)
# ----------------------------------------------------------------------------------------
def cg(A_mm: Callable[[np.ndarray], np.ndarray], b: np.ndarray, rtol=1e-3, atol=0.0, maxiter=None,
       x0: Optional[np.ndarray] = None, fixed_iters: Optional[int] = None):
    """x0 defaults to b (M_mm = identity, cg.py:141-144).  ``fixed_iters`` (oracle-only test hook)
    disables the residual test and runs exactly that many iterations."""
    n = b.shape[0]
    if maxiter is None:
        maxiter = 5 * n
    x = b.copy() if x0 is None else x0.copy()
    r = b - A_mm(x)
    p = r.copy()
    r_norm = np.linalg.norm(r)
    stop = max(rtol * np.linalg.norm(b), atol)
    optimal = False
    k = 0
    for k in range(1, maxiter + 1):
        Ap = A_mm(p)
        pAp = p @ Ap
        if pAp <= 1e-16:                                   # :250
            break
        rr = r @ r
        alpha = rr / pAp
        x = x + alpha * p
        r_new = r - alpha * Ap
        r_norm_new = np.linalg.norm(r_new)
        if fixed_iters is None and r_norm_new <= stop:     # :259
            optimal = True
            r, r_norm = r_new, r_norm_new
            break
        beta = (r_new @ r_new) / rr
        p = r_new + beta * p
        r, r_norm = r_new, r_norm_new
        if fixed_iters is not None and k >= fixed_iters:
            break
    return x, {"niter": k, "optimal": optimal, "residual_norm": r_norm}


def rtol_func(sigma, rtol_max=1e0, rtol_min=1e-14):
    """conditioning_mechanisms.py:307-323 (the clamp expression at :313 is reproduced verbatim)."""
    sigma_min, sigma_max = 0.1, 80.0
    sigma = max(min(sigma, sigma_max), max(sigma_min, sigma))
    p = 0.1
    log_factor = ((math.log10(sigma) - math.log10(sigma_min)) / (math.log10(sigma_max) - math.log10(sigma_min))) ** p
    log_rtol = log_factor * (math.log10(rtol_max) - math.log10(rtol_min)) + math.log10(rtol_min)
    return 10 ** log_rtol


# ----------------------------------------------------------------------------------------
# Operators  (measurement_utils/measurements.py:87-320, measurement_utils/utils_sisr.py:22-96)
# ----------------------------------------------------------------------------------------
def p2o(psf: np.ndarray, shape: Tuple[int, int]) -> np.ndarray:
    """PSF -> OTF.  utils_sisr.py:22-41: zero-pad to `shape`, roll by -int(ks/2) on both axes, fft2.
    The reference does this in float32 (torch.Tensor kernels): the pad/roll is exact, the FFT is complex64."""
    otf = np.zeros(psf.shape[:-2] + tuple(shape), dtype=psf.dtype)
    otf[..., :psf.shape[-2], :psf.shape[-1]] = psf
    for axis, axis_size in enumerate(psf.shape[-2:]):
        otf = np.roll(otf, -int(axis_size / 2), axis=otf.ndim - 2 + axis)
    return sfft.fft2(otf, axes=(-2, -1))


def upsample(x: np.ndarray, sf: int) -> np.ndarray:
    """utils_sisr.py:44-52: zero insertion, samples at [0::sf, 0::sf]."""
    z = np.zeros(x.shape[:-2] + (x.shape[-2] * sf, x.shape[-1] * sf), dtype=x.dtype)
    z[..., 0::sf, 0::sf] = x
    return z


def downsample(x: np.ndarray, sf: int) -> np.ndarray:
    """utils_sisr.py:55-61."""
    return x[..., 0::sf, 0::sf]


class _FFTOperator:
    """Circular convolution through the OTF: forward = ifft2(FB fft2 x).real, transpose uses conj(FB).
    measurements.py:140-157 (motion), :180-198 (gaussian); pre_calculate utils_sisr.py:79-96."""
    scale_factor = 1

    def __init__(self, kernel: np.ndarray, sigma_s: float, image_size: int = 256, otf_dtype=np.complex64):
        self.kernel = np.asarray(kernel, np.float32)            # torch.Tensor(kernel): float32 (measurements.py:160,175)
        self.sigma_s = float(np.float32(sigma_s))
        self.image_size = image_size
        FB = p2o(self.kernel[None, None], (image_size * self.scale_factor,) * 2).astype(otf_dtype)
        self.FB, self.FBC = FB, np.conj(FB)

    def A(self, x):
        return sfft.ifft2(self.FB * sfft.fft2(x, axes=(-2, -1)), axes=(-2, -1)).real

    def AT(self, y):
        return sfft.ifft2(self.FBC * sfft.fft2(y, axes=(-2, -1)), axes=(-2, -1)).real

    def forward(self, x, noise=None):
        y = self.A(x)
        if noise is not None:
            y = y + self.sigma_s * noise
        return y


class GaussianBlurOracle(_FFTOperator):
    name = "gaussian_blur"


class MotionBlurOracle(_FFTOperator):
    name = "motion_blur"


class SuperResolutionOracle(_FFTOperator):
    """Solver-side SR operator A = downsample_sf . (k (*)), A^T = (k (*))^T . upsample_sf.
    conditioning_mechanisms.py:641-675; kernel = kernels[0, sf-2] of kernels_bicubicx234.mat (measurements.py:96-98).
    (The reference synthesises y with a different, bicubic `Resizer` -- measurements.py:105 -- which is a
    once-per-image step outside this oracle; y is an input here.)"""
    name = "super_resolution"

    def __init__(self, kernel, sigma_s, scale_factor=4, image_size=256, otf_dtype=np.complex64):
        self.scale_factor = 1
        super().__init__(kernel, sigma_s, image_size, otf_dtype)
        self.scale_factor = scale_factor

    def A(self, x):
        return downsample(super().A(x), self.scale_factor)

    def AT(self, y):
        return super().AT(upsample(y, self.scale_factor))


# ---- measurement synthesis for super-resolution: the bicubic `Resizer` --------------------------
def _cubic(t):
    """Keys cubic (a = -0.5), measurement_utils/resizer.py:171-176."""
    a = np.abs(t)
    return np.where(a <= 1, (1.5 * a - 2.5) * a * a + 1, np.where(a <= 2, ((-0.5 * a + 2.5) * a - 4) * a + 2, 0.0))


def resizer_matrix(n_in: int, scale: float, antialiasing: bool = True) -> np.ndarray:
    """Dense (n_out, n_in) matrix of one dimension of Resizer(kernel='cubic') -- resizer.py:104-165: output pixel u
    (1-based) sits at c = u'/scale + (1 - 1/scale)/2 with u' = u - (n_out - n_in*scale)/2 (:117-130); with antialiasing
    (downscaling only, :27) the kernel is scale*cubic(scale*t), 4/scale wide (:112-113); ceil(width)+2 candidate taps
    start at floor(c - width/2) - 1 (:133-143), weights are normalised to sum 1 (:151-153) and indices are reflected at
    the borders (:156-157).  Weights are rounded to float32 like the reference's weight tensors (:44)."""
    n_out = int(np.ceil(n_in * scale))                                    # :98
    aa = antialiasing and scale < 1
    width = 4.0 / scale if aa else 4.0
    kern = (lambda t: scale * _cubic(scale * t)) if aa else _cubic
    u = np.arange(1, n_out + 1) - (n_out - n_in * scale) / 2.0
    c = u / scale + 0.5 * (1.0 - 1.0 / scale)
    left = np.floor(c - width / 2.0)
    taps = int(np.ceil(width)) + 2
    idx = (left[:, None] + np.arange(taps)[None, :] - 1).astype(np.int64)          # 0-based input index, may be out of range
    w = kern(c[:, None] - idx - 1.0)
    tot = w.sum(axis=1)
    tot[tot == 0] = 1.0
    w = (w / tot[:, None]).astype(np.float32).astype(np.float64)
    idx = np.mod(idx, 2 * n_in)
    idx = np.where(idx >= n_in, 2 * n_in - 1 - idx, idx)                             # mirror
    R = np.zeros((n_out, n_in))
    np.add.at(R, (np.repeat(np.arange(n_out), taps), idx.reshape(-1)), w.reshape(-1))
    return R


def bicubic_downsample(x, scale_factor: int):
    """y = Resizer(in_shape, 1/scale_factor)(x) over the last two axes (measurements.py:91,105): the measurement the
    reference synthesises for super-resolution (NOT the operator A its solver uses)."""
    x = np.asarray(x, np.float64)
    Rh = resizer_matrix(x.shape[-2], 1.0 / scale_factor)
    Rw = resizer_matrix(x.shape[-1], 1.0 / scale_factor)
    return np.einsum("oh,...hw,pw->...op", Rh, x, Rw)


def mask_random(image_size: int, prob_range: Tuple[float, float], rng=np.random) -> np.ndarray:
    """measurements.py:287-299.  Consumes the *global* numpy RNG exactly like the reference:
    one uniform() then one choice(replace=False).  Returns (1,3,H,W) float32 of {0,1}, 1 = observed."""
    total = image_size ** 2
    l, h = prob_range
    prob = rng.uniform(l, h)
    mask_vec = np.ones(total, np.float32)
    samples = rng.choice(total, int(total * prob), replace=False)
    mask_vec[samples] = 0
    m = mask_vec.reshape(1, image_size, image_size)
    return np.broadcast_to(m, (3, image_size, image_size))[None].copy()


def mask_box(image_size: int, len_range: Tuple[int, int], margin=(16, 16), rng=np.random) -> np.ndarray:
    """measurements.py:276-285,301-320: h,w = randint(l,h) each; t,l = randint(margin, size-margin-len)."""
    l, h = int(len_range[0]), int(len_range[1])
    mh = rng.randint(l, h)
    mw = rng.randint(l, h)
    maxt = image_size - margin[0] - mh
    maxl = image_size - margin[1] - mw
    t = rng.randint(margin[0], maxt)
    le = rng.randint(margin[1], maxl)
    m = np.ones((1, 3, image_size, image_size), np.float32)
    m[..., t:t + mh, le:le + mw] = 0
    return m


class InpaintingOracle:
    """measurements.py:204-246: forward = (x + sigma_s n) * mask; A = A^T = mask."""
    name = "inpainting"
    scale_factor = 1

    def __init__(self, mask: np.ndarray, sigma_s: float):
        self.mask = np.asarray(mask, np.float32)
        self.sigma_s = float(np.float32(sigma_s))

    def A(self, x):
        return self.mask * x

    AT = A

    def forward(self, x, noise=None):
        y = x if noise is None else x + self.sigma_s * noise
        return y * self.mask


# ----------------------------------------------------------------------------------------
# mat solvers  (conditioning_mechanisms.py:384-419, 489-527, 641-675; dispatch :562-600)
# ----------------------------------------------------------------------------------------
def mat_solver(operator, y, x0_mean, covariance_model, sigma_t, max_rtol=1.0, maxiter=5000,
               fixed_iters: Optional[int] = None):
    """Solves (A Sigma A^T + sigma_y^2 I) u = y - A x0_mean with CG (x0 = b) and returns (A^T u, info).
    sigma_y = sigma_s.clip(min=1e-3) (:387,:491), additionally clip(min=1e-2) for SR (:642)."""
    y = np.asarray(y, np.float64)
    mu = np.asarray(x0_mean, np.float64)
    # sigma_s is a 1-element float32 tensor in the reference; clip and square happen in float32
    sig = np.maximum(np.float32(operator.sigma_s), np.float32(1e-3))
    if operator.name == "super_resolution":
        sig = np.maximum(sig, np.float32(1e-2))
    sig2 = float(np.float32(sig * sig))
    shape_y = y.shape

    def A_mm(u):
        u = u.reshape(shape_y)
        out = sig2 * u + operator.A(covariance_model.denoiser_cov_vector_dot(operator.AT(u)))
        return out.reshape(-1)

    if operator.name == "inpainting":
        b = operator.mask * y - operator.mask * mu                   # :407
    else:
        b = y - operator.A(mu)                                       # :515 / :663
    u, info = cg(A_mm, b.reshape(-1), rtol=rtol_func(sigma_t, max_rtol), maxiter=maxiter, fixed_iters=fixed_iters)
    u = u.reshape(shape_y)
    mat = u if operator.name == "inpainting" else operator.AT(u)     # :417 / :525 / :673
    return mat, info


# ----------------------------------------------------------------------------------------
# BFGSOnlineUpdate  (conditioning_mechanisms.py:190-294) -- the Free Hunch mechanism
# ----------------------------------------------------------------------------------------
def _allclose(a, b, rtol=1e-5, atol=1e-8):
    """torch.allclose(a, b) semantics: |a-b| <= atol + rtol*|b| everywhere."""
    return bool(np.all(np.abs(a - b) <= atol + rtol * np.abs(b)))


@dataclass
class FreeHunchOracle:
    """`model(x, sigma) -> (x0_mean, vjp)` where vjp(w) = J^T w is the vector-Jacobian product of the
    denoiser mean at x (the reference gets it from autograd, conditioning_mechanisms.py:280)."""
    operator: object
    covariance_model: CovarianceOracle
    cond_scaling: float = 1.0
    clip_x0_mean: bool = False
    do_space_updates: bool = True
    space_step_update_threshold: float = 10.0
    space_step_update_lower_threshold: float = 1.0
    denoiser_mean_error_threshold: float = 0.2
    max_rtol: float = 1.0
    cg_maxiter: int = 5000
    sigmas: List[float] = field(default_factory=list)
    xs: List[np.ndarray] = field(default_factory=list)
    denoiser_means: List[np.ndarray] = field(default_factory=list)
    log: List[dict] = field(default_factory=list)

    def __call__(self, x_t, model, y, sigma):                      # :46-50
        out = self.x0_mean_update(x_t, model, y, sigma)
        return np.clip(out, -1, 1) if self.clip_x0_mean else out

    def x0_mean_update(self, x_t, model, y, sigma):                # :237-294
        sigma = float(sigma)
        x_t = np.asarray(x_t, np.float64)
        mu, vjp = model(x_t, sigma)
        rec = {"sigma": sigma, "time_update": False, "space_update": False}
        cov = self.covariance_model
        if self.do_space_updates:
            mu_prev_at_sigma = None
            if self.sigmas and sigma != self.sigmas[-1]:           # :241-244
                score_prev = (self.denoiser_means[-1] - self.xs[-1]) / self.sigmas[-1] ** 2
                mu_prev_at_sigma, _ = cov.update_time_step(self.xs[-1], self.sigmas[-1], sigma, score_prev)
                rec["time_update"] = True
            elif self.sigmas:                                      # :248-249
                mu_prev_at_sigma = self.denoiser_means[-1]
            if self.xs and not _allclose(x_t, self.xs[-1]):        # :250
                if self.space_step_update_lower_threshold < sigma < self.space_step_update_threshold:   # :256
                    cov.update_space_step(mu_prev_at_sigma, mu, sigma, self.xs[-1], x_t)
                    rec["space_update"] = True
        elif self.sigmas and sigma != self.sigmas[-1]:             # :261-263
            score_prev = (self.denoiser_means[-1] - self.xs[-1]) / self.sigmas[-1] ** 2
            cov.update_time_step(self.xs[-1], self.sigmas[-1], sigma, score_prev, only_covariance=True)
            rec["time_update"] = True

        mat, info = mat_solver(self.operator, y, mu, cov, sigma, self.max_rtol, self.cg_maxiter)   # :268
        g = vjp(mat)                                               # :280
        if np.std(g * sigma ** 2, ddof=1) > self.denoiser_mean_error_threshold:                  # :283
            g = cov.denoiser_cov_vector_dot(mat) * self.cond_scaling / sigma ** 2               # :284
            rec["fallback"] = True
        else:
            g = g * self.cond_scaling                              # :287
            rec["fallback"] = False
        out = mu + g * sigma ** 2                                  # :285/:288
        self.sigmas.append(sigma)
        self.xs.append(x_t.copy())
        self.denoiser_means.append(mu.copy())
        rec["niter"] = info["niter"]
        rec["k"] = cov.cov_u.shape[-1]
        self.log.append(rec)
        return out


# ----------------------------------------------------------------------------------------
# Sampler  (generate_conditional.py:110-167,172-201; training/openai_preconditioning.py:116-129,203-207)
# ----------------------------------------------------------------------------------------
def iddpm_linear_u(M=1000, beta_min=1e-4, beta_max=0.02) -> np.ndarray:
    """openai_preconditioning.py:116-122: u_j for betas=[0, linspace(beta_min,beta_max,M)] (float32 like torch)."""
    betas = np.concatenate([[0.0], np.linspace(beta_min, beta_max, M, dtype=np.float64)]).astype(np.float32)
    alphas = (1 - betas).astype(np.float32)
    alpha_bar = np.cumprod(alphas, dtype=np.float32)[::-1]
    return np.sqrt((1 - alpha_bar) / alpha_bar).astype(np.float32)


def round_sigma(sigma: np.ndarray, u: np.ndarray) -> np.ndarray:
    """openai_preconditioning.py:203-207: nearest u_j in float32, returned in sigma's dtype."""
    s32 = np.asarray(sigma, np.float32).reshape(-1)
    idx = np.abs(s32[:, None] - u[None, :]).argmin(1)
    return u[idx].astype(np.asarray(sigma).dtype).reshape(np.shape(sigma))


def edm_sigma_steps(num_steps, sigma_min=0.002, sigma_max=80.0, rho=7.0, u: Optional[np.ndarray] = None) -> np.ndarray:
    """generate_conditional.py:74-75,82,200,111-112: EDM grid clipped to the net range, rounded, 0 appended (float64)."""
    if u is None:
        u = iddpm_linear_u()
    net_min, net_max = float(u[len(u) - 2]), float(u[0])
    sigma_min, sigma_max = max(sigma_min, net_min), min(sigma_max, net_max)
    i = np.arange(num_steps, dtype=np.float64)
    steps = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return np.concatenate([round_sigma(steps, u), [0.0]])


def conditional_sampler(mechanism, model, y, noise, t_steps, solver="heun"):
    """generate_conditional.py:114-167 with schedule='linear', scaling='none', S_churn=0, alpha=1:
    x0 = noise*t0; per step Euler with the guided denoiser, Heun correction on all but the last step."""
    x_next = np.asarray(noise, np.float64) * t_steps[0]
    n = len(t_steps) - 1
    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_hat = x_next
        h = t_next - t_cur
        den = mechanism(x_hat, model, y, t_cur)
        d_cur = (x_hat - den) / t_cur                       # :146-147
        x_prime = x_hat + h * d_cur                         # :149
        if solver == "euler" or i == n - 1:
            x_next = x_hat + h * d_cur                      # :154
        else:
            den2 = mechanism(x_prime, model, y, t_next)     # :160
            d_prime = x_prime / t_next - den2 / t_next      # :166
            x_next = x_hat + h * (0.5 * d_cur + 0.5 * d_prime)   # :167
    return x_next


def analytic_dct_gaussian_denoiser(dct_variance: np.ndarray):
    """Stand-in denoiser used by the benchmarks and tests: exact for a
    DCT-diagonal Gaussian prior, linear and symmetric, so its VJP is the same linear map."""
    v = np.asarray(dct_variance, np.float64)

    def model(x, sigma):
        gain = v / (v + sigma ** 2)
        lin = lambda w: idct2(gain * dct2(w))
        return lin(x), lin
    return model


# ----------------------------------------------------------------------------------------
# Gaussian-mixture toy (config 1): notebooks/diffusion_gaussian_mixture_utils.py
# ----------------------------------------------------------------------------------------
def gmm_score(x, weights, means, covs, sigma):
    """grad_x log p_sigma(x) for a Gaussian mixture.  diffusion_gaussian_mixture_utils.py:165-206."""
    d = x.shape[1]
    comb = covs + sigma ** 2 * np.eye(d)[None]
    prec = np.linalg.inv(comb)
    diff = means[None] - x[:, None]                                   # (bs,n,d)
    quad = np.einsum("bni,nij,bnj->bn", diff, prec, diff)
    logp = -0.5 * quad + np.log(weights)[None]
    logp = logp - np.log(np.sum(np.exp(logp - logp.max(1, keepdims=True)), 1, keepdims=True)) - logp.max(1, keepdims=True)
    resp = np.exp(logp)
    return np.einsum("bn,nij,bnj->bi", resp, prec, diff)


def gmm_score_jacobian(x, weights, means, covs, sigma):
    """Exact Hessian of log p_sigma (what the notebook gets by autograd, :286-310), closed form."""
    d = x.shape[1]
    comb = covs + sigma ** 2 * np.eye(d)[None]
    prec = np.linalg.inv(comb)
    diff = means[None] - x[:, None]
    quad = np.einsum("bni,nij,bnj->bn", diff, prec, diff)
    logdet = np.linalg.slogdet(comb)[1]
    logp = -0.5 * quad + np.log(weights)[None]
    # log-responsibilities: the notebook omits the log-determinants (:193-194), so do we
    logp = logp - logp.max(1, keepdims=True)
    resp = np.exp(logp) / np.exp(logp).sum(1, keepdims=True)
    g = np.einsum("nij,bnj->bni", prec, diff)                         # per-component score (bs,n,d)
    s = np.einsum("bn,bni->bi", resp, g)
    H = -np.einsum("bn,nij->bij", resp, prec) + np.einsum("bn,bni,bnj->bij", resp, g, g) - np.einsum("bi,bj->bij", s, s)
    del logdet
    return H


def gmm_covariance(weights, means, covs):
    """Overall covariance of the mixture.  diffusion_gaussian_mixture_utils.py:815-837."""
    mean = (weights[:, None] * means).sum(0)
    c = np.zeros((means.shape[1],) * 2)
    for w, m, s in zip(weights, means, covs):
        c += w * s + w * np.outer(m - mean, m - mean)
    return c


def gmm_guidance(x, y, weights, means, covs, sigma, cov_y, Sigma):
    """grad = J_mu^T (Sigma_y + Sigma)^-1 (y - mu(x)).  diffusion_gaussian_mixture_utils.py:663-702
    (J_mu = I + sigma^2 * Hessian, symmetric)."""
    mu = x + sigma ** 2 * gmm_score(x, weights, means, covs, sigma)
    A = np.linalg.inv(cov_y[None] + Sigma)
    v = np.einsum("bij,bj->bi", A, y[None] - mu)
    J = np.eye(x.shape[1])[None] + sigma ** 2 * gmm_score_jacobian(x, weights, means, covs, sigma)
    return np.einsum("bji,bj->bi", J, v)


def gmm_sample_bfgs(prior_samples, sigma_of_t, dsigma_dt, num_steps, init_cov, y, weights, means, covs, cov_y, rho=1.0):
    """Euler ODE sampler with dense time+BFGS covariance tracking.
    diffusion_gaussian_mixture_utils.py:926-987 (reverse_ode_step_euler: x - dt * (-sigma' sigma score))."""
    bs, d = prior_samples.shape
    eye = np.eye(d)
    cov = np.repeat(init_cov[None], bs, 0)
    inv_cov = np.linalg.inv(cov)
    s1 = sigma_of_t(1.0)
    hes = (cov / s1 ** 2 - eye[None]) / s1 ** 2
    inv_hes = np.linalg.inv(hes)
    ts = np.linspace(1, 0, num_steps + 1) ** rho
    dts = ts[:-1] - ts[1:]
    x = prior_samples.copy()
    score = gmm_score(x, weights, means, covs, sigma_of_t(ts[0]))
    upd = score + gmm_guidance(x, y, weights, means, covs, sigma_of_t(ts[0]), cov_y, cov)
    covs_out = [cov]
    for i, t in enumerate(ts[:-1]):
        dt = dts[i]
        sig, sig_n = sigma_of_t(t), sigma_of_t(t - dt)
        xnew = x + dt * dsigma_dt(t) * sig * upd
        dx = xnew - x
        score_new = gmm_score(xnew, weights, means, covs, sig_n)
        mean_new = xnew + sig_n ** 2 * score_new
        cov, inv_cov, hes, inv_hes, _, _ = dense_update_covariance(x, cov, inv_cov, hes, inv_hes, score, sig, sig_n)
        mean_at_tnext = x + sig_n ** 2 * gmm_score(x, weights, means, covs, sig_n)     # :966
        cov, inv_cov, hes, inv_hes = dense_update_bfgs(cov, inv_cov, mean_at_tnext, mean_new, sig_n, dx)
        covs_out.append(cov)
        score = score_new
        upd = score + gmm_guidance(xnew, y, weights, means, covs, sig_n, cov_y, cov)
        x = xnew
    return x, covs_out
