"""Host-side mirror of the reference covariance models (conditioning_utils/online_update_bfgs.py).

Same class names, constructor arguments, method names and error behaviour as the reference so its tests
read unchanged, but every method is one call into the sm_100a library: the state lives in HBM as the prior
precision diag(pi) + Z^T Gz Z, nothing is computed on the host and nothing is copied to the host.
Beyond the reference: `batch` > 1 keeps B independent per-image states on one GPU (the reference asserts
batch size 1, online_update_bfgs.py:161,255).
"""
import os

import numpy as np
import torch

from . import engine

_CTX_CACHE = {}


def get_context(device, channels, height, width, max_batch=1):
    """One workspace per (device, geometry); grown when a larger batch is requested."""
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (str(device), channels, height, width)
    ctx = _CTX_CACHE.get(key)
    if ctx is None or ctx.max_batch < max_batch:
        ctx = engine.Context(device, max_batch=max_batch, channels=channels, height=height, width=width)
        _CTX_CACHE[key] = ctx
    return ctx


def _default_device():
    return torch.device("cuda", torch.cuda.current_device())


class CovarianceHessianBFGS:
    """Identity-basis model.  Reference: online_update_bfgs.py:7-336."""
    _basis = engine.BASIS_IDENTITY

    def __init__(self, init_denoiser_variance, init_noise_variance, data_dim, dtype=None, max_vector_count=None,
                 init_denoiser_cov_u=None, project_to_diagonal=False, use_precalculated_info=True, device=None,
                 batch=1, image_shape=None):
        self.data_dim = int(data_dim)
        self.dtype = dtype
        self.max_vector_count = max_vector_count
        self.project_to_diagonal = project_to_diagonal
        self.batch = batch
        device = torch.device(device) if device is not None else _default_device()
        if image_shape is None:
            image_shape = (1, 1, self.data_dim)
        C, H, W = image_shape
        assert C * H * W == self.data_dim, "image_shape does not match data_dim"
        self.image_shape = (C, H, W)
        self.ctx = get_context(device, C, H, W, max_batch=batch)
        self.state = engine.CovState(self.ctx, batch=batch, basis=self._basis)
        var = torch.as_tensor(init_denoiser_variance, dtype=torch.float32).reshape(-1)
        if var.numel() == 1:
            var = var.expand(self.data_dim)
        self._sigma = float(np.sqrt(float(init_noise_variance)))
        self.state.init(var, float(init_noise_variance))
        self._pending_u = self._pending_v = None
        # factor history for drop_vectors(0 < n < k): only when a finite max_vector_count can ever trigger it
        self._trackers = None
        if max_vector_count is not None and 0 < max_vector_count < engine.MAX_RANK // 2 and not project_to_diagonal:
            from .ref_factors import ReferenceFactorTracker
            self._trackers = [ReferenceFactorTracker() for _ in range(batch)]
        if init_denoiser_cov_u is not None:                     # online_update_bfgs.py:16-18,37
            self.vectors_denoiser_cov_u = init_denoiser_cov_u
            self.set_others_corresponding_to_current_denoiser_cov(self._sigma)

    # ---- shapes -------------------------------------------------------------------------------
    def _to_batch(self, t):
        t = torch.as_tensor(t)
        return t.reshape(self.batch, *self.image_shape)

    def _check_batch(self, shape):
        # the reference asserts shape[0] == 1 (online_update_bfgs.py:161,255); here the leading dim is the batch
        assert len(shape) >= 1 and shape[0] == self.batch, f"Batch size must be {self.batch}"

    # ---- reference API ------------------------------------------------------------------------
    def update_time_step(self, x_t, sigma_t, sigma_tnext, score_t, only_covariance=False):
        shape, dtype = tuple(x_t.shape), x_t.dtype
        self._check_batch(shape)
        self._sigma = float(sigma_tnext)
        if only_covariance:
            self.state.time_update(None, None, sigma_t, sigma_tnext, only_covariance=True)
            self._track("time_update")
            x = x_t.detach()
            return x.clone(), x.clone()                         # online_update_bfgs.py:187-188
        mean, score = self.state.time_update(self._to_batch(x_t.detach()), self._to_batch(score_t.detach()),
                                             float(sigma_t), float(sigma_tnext))
        self._track("time_update")
        return mean.reshape(shape).to(dtype), score.reshape(shape).to(dtype)

    def update_space_step(self, denoiser_mean_at_x, denoiser_mean_at_xnext, sigma_t, x, xnext):
        self._check_batch(tuple(x.shape))
        self._sigma = float(sigma_t)
        self.state.space_update(self._to_batch(denoiser_mean_at_x.detach()), self._to_batch(denoiser_mean_at_xnext.detach()),
                                float(sigma_t), self._to_batch(x.detach()), self._to_batch(xnext.detach()),
                                project_to_diagonal=self.project_to_diagonal)
        if not self.project_to_diagonal:
            self._track("space_update")
        if self.max_vector_count is not None:                   # online_update_bfgs.py:309-310
            self.drop_vectors(self.max_vector_count, sigma_t)

    def denoiser_cov_vector_dot(self, v, use_cuda=True):
        shape, dtype = tuple(v.shape), v.dtype
        out = self.state.cov_dot(self._to_batch(v.detach()))
        return out.reshape(shape).to(dtype)

    def drop_vectors(self, max_vector_count, sigma):
        """online_update_bfgs.py:233-245.  n = 0 forgets the low-rank part.  0 < n < k keeps "the last n columns" of
        the reference's factors U, V -- columns of its own Woodbury / sqrtm re-factorisations (:87-119), which the product
        does not store.  With `max_vector_count` set, `ref_factors.ReferenceFactorTracker` carries their coefficients over
        the rows of Z (one m x m Gram read per update), and a drop moves the state to the basis of the kept columns
        (`fh_state_recombine`), which bounds the rank at 2 n + 2 like the reference's memory.  Refused (NotImplementedError)
        where the kept columns are complex: the reference then continues with a complex 'covariance' and takes real
        parts of its products, which no real-valued state reproduces."""
        k = self.state.rank // 2
        if max_vector_count == 0:
            self.state.drop_all(float(sigma))
            for t in self._trackers or ():
                t.reset()
        elif k > max_vector_count:
            if self._trackers is None:
                raise NotImplementedError("drop_vectors(0 < n < rank) needs the factor history: construct the model with "
                                          "max_vector_count set")
            K = self.state.gram().cpu().numpy()
            drops = [t.drop(int(max_vector_count), K[b]) for b, t in enumerate(self._trackers)]
            self.state.recombine(np.stack([d[0] for d in drops]), np.stack([d[1] for d in drops]))

    def _track(self, what):
        """Keeps the reference's factor coefficients in step with the state (only when max_vector_count is set)."""
        if self._trackers is None or self.state.rank == 0:
            return
        K = self.state.gram().cpu().numpy()
        for b, t in enumerate(self._trackers):
            getattr(t, what)(K[b])

    @property
    def vector_count(self):
        return self.state.rank // 2

    # ---- test hooks used by the reference tests (online_update_tests.py:153-155) ------------------
    @property
    def vectors_denoiser_cov_u(self):
        return self._cov_factors()[0]

    @vectors_denoiser_cov_u.setter
    def vectors_denoiser_cov_u(self, U):
        self._pending_u = torch.as_tensor(U)

    @property
    def vectors_denoiser_cov_v(self):
        return self._cov_factors()[1]

    @vectors_denoiser_cov_v.setter
    def vectors_denoiser_cov_v(self, V):
        self._pending_v = torch.as_tensor(V)

    def set_others_corresponding_to_current_denoiser_cov(self, sigma):
        """online_update_bfgs.py:327-330: rebuild everything from Sigma = diag + U U^T - V V^T."""
        d = self.data_dim
        U = self._pending_u if self._pending_u is not None else torch.zeros(d, 0)
        V = self._pending_v if self._pending_v is not None else torch.zeros(d, 0)
        U = U.real if U.is_complex() else U
        V = V.real if V.is_complex() else V
        self.state.set_cov_vectors(U.to(torch.float32).T.contiguous(), V.to(torch.float32).T.contiguous(), float(sigma))
        self._sigma = float(sigma)
        self._pending_u = self._pending_v = None

    def _cov_factors(self):
        """Real (U, V) with Sigma = diag(delta) + U U^T - V V^T for image 0 (eigen-split of the core)."""
        pi, Z, Gz, Gs, delta = self.state.export()
        m = Z.shape[1]
        if m == 0:
            z = torch.zeros(self.data_dim, 0, dtype=torch.float64, device=Z.device)
            return z, z.clone()
        lam, Q = torch.linalg.eigh(Gs[0])
        Y = (Z[0].double() * delta[0].double()).T                 # d x m
        pos, neg = lam > 0, lam < 0
        U = Y @ (Q[:, pos] * lam[pos].sqrt())
        V = Y @ (Q[:, neg] * (-lam[neg]).sqrt())
        return U, V

    def get_dense_matrices(self):
        """(Sigma, Sigma^-1, H, H^-1) of image 0 as dense float64 matrices (online_update_bfgs.py:320-325).
        Test hook for small d: the dense matrices are formed from the exported state with torch."""
        pi, Z, Gz, Gs, delta = self.state.export()
        d = self.data_dim
        sigma2 = self.state.var2
        eye = torch.eye(d, dtype=torch.float64, device=pi.device)
        inv_cov = torch.diag(pi[0].double() + 1.0 / sigma2)
        if Z.shape[1]:
            inv_cov = inv_cov + Z[0].double().T @ Gz[0] @ Z[0].double()
        cov = torch.linalg.inv(inv_cov)
        hes = (cov / sigma2 - eye) / sigma2
        return cov, inv_cov, hes, torch.linalg.inv(hes)

    # basis hooks (online_update_bfgs.py:314-318)
    def transform(self, x):
        return x

    def inverse_transform(self, x):
        return x


class CovarianceHessianBFGSDCT(CovarianceHessianBFGS):
    """All operations in the per-channel orthonormal 2-D DCT basis.  Reference: online_update_bfgs.py:339-374.
    `data_dir` must contain dct_variance.pt (3,256,256) exactly like the reference (:343); a tensor can be
    passed instead through `dct_variance=`."""
    _basis = engine.BASIS_DCT

    def __init__(self, data_dir, init_noise_variance, data_dim, dtype=None, max_vector_count=None,
                 dct_variance=None, use_precalculated_info=True, project_to_diagonal=False, device=None, batch=1,
                 image_shape=(3, 256, 256), **kwargs):
        if dct_variance is None:
            dct_variance = torch.load(os.path.join(data_dir, "dct_variance.pt"))
        dct_variance = torch.as_tensor(dct_variance, dtype=torch.float32)
        self.dct_variance = dct_variance if use_precalculated_info else torch.ones(data_dim)        # :345-348
        super().__init__(self.dct_variance.reshape(-1), init_noise_variance, data_dim, dtype, max_vector_count,
                         project_to_diagonal=project_to_diagonal, device=device, batch=batch, image_shape=image_shape)

    def transform(self, x):
        return self.ctx.dct2(x)

    def inverse_transform(self, x):
        return self.ctx.dct2(x, inverse=True)


class DCTVarianceAccumulator:
    """The prior the DCT-basis model starts from: per-coefficient variance of the orthonormal 2-D DCT over a set of images
    (do_frequency_analysis.py:26-72, which writes `dct_variance.pt`).  Batches of (N, 3, 256, 256) images in [-1, 1] go through
    `fh_dct2` on the device; the sums of the coefficients and of their squares are kept on the device in float64 (the
    reference sums in float32 on the CPU; over its 500 images the two agree to float32 rounding).
    variance() = E[c^2] - E[c]^2 as float32, `save(data_dir)` writes the file CovarianceHessianBFGSDCT loads."""

    def __init__(self, ctx, image_shape=(3, 256, 256)):
        self.ctx = ctx
        self.sum = torch.zeros(image_shape, dtype=torch.float64, device=ctx.device)
        self.sum_sq = torch.zeros_like(self.sum)
        self.count = 0

    def update(self, images, limit=None):
        """Adds a batch; with `limit`, only as many images as keep the total at or below it (:36-39).  Returns the total."""
        if limit is not None:
            images = images[: max(0, int(limit) - self.count)]
        if images.shape[0] == 0:
            return self.count
        c = self.ctx.dct2(images.to(torch.float32)).to(torch.float64)
        self.sum += c.sum(dim=0)
        self.sum_sq += (c * c).sum(dim=0)
        self.count += int(images.shape[0])
        return self.count

    def variance(self):
        if self.count == 0:
            raise ValueError("no images accumulated")
        mean = self.sum / self.count
        return (self.sum_sq / self.count - mean * mean).to(torch.float32)

    def save(self, data_dir):
        os.makedirs(data_dir, exist_ok=True)
        path = os.path.join(data_dir, "dct_variance.pt")
        torch.save(self.variance().cpu(), path)
        return path
